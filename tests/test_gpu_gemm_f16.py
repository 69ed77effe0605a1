"""The default path for every n the GEMV does not take (n > 8, or fewer columns of a long row: b200_gemm_f16.cu): weights dequantized to fp16 inside the kernel, Q8_0-quantized activations
as fp16, ONE dense contraction with fp32 accumulation on the tensor cores (tcgen05 cta_group::2, a CTA pair per 256 x 256 tile,
persistent, tail tiles split along k) -- what the reference's CUDA backend does for large batches (ggml_cuda_op_mul_mat_cublas,
src/ggml-cuda.cu:1208-1306).  Bounds: NMSE <= 5e-4 against the oracle is the reference's bar (tests/test-backend-ops.cpp:921-923);
one fp16 rounding per operand element (relative 2^-12) leaves it around 1e-7: the tests hold it to 1e-6, against the oracle AND
against the exact int8 kernel."""
import numpy as np
import pytest

from conftest import Q4_0, Q8_0, MUL_MAT_NMSE_TOL, F16_GEMM_NMSE, nmse

pytestmark = pytest.mark.gpu

SHAPES = [
    (128, 64, 256), (128, 256, 256), (300, 256, 64), (1000, 4096, 512), (257, 96, 33), (64, 32, 300),   # ragged m / n / k % 64 == 32
    (256, 64, 256), (512, 2048, 512),              # exactly one / a few whole tiles
    (4096, 4096, 512),                             # 32 tiles on 74 pairs: every tile split along k
    (11008, 4096, 512),                            # BASELINE C2: one full round + 12 tiles split 6 ways
    (16384, 4096, 512), (4096, 16384, 128),        # GPT-J fc_in / fc_out
    (20000, 512, 700),                             # several full rounds, ragged n, short k (no split possible)
    (2304, 768, 128), (50257 // 8, 768, 128),      # GPT-2 qkv and (a slice of) lm_head at the 128-token prompt
    (4096, 4096, 9), (300, 256, 17), (4096, 16384, 31), (1000, 96, 12),      # short prompts: a few columns of one padded tile
    (4096, 16384, 4), (16384, 4096, 8),            # the columns at which the GEMV's activations stop fitting beside its weight ring (Q8_0: also 16384 x 4096 x 8)
]


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("m,k,n", SHAPES)
def test_f16_gemm_vs_oracle_and_exact_kernel(gpu_ctx, qmm, oracle, qtype, m, k, n):
    rng = np.random.default_rng(m * 31 + k + n)
    big = m * k > (1 << 24)
    wire = qmm.random_wire_weights(qtype, k, m, seed=m + k) if big else oracle.quantize_weights(qtype, rng.uniform(-1, 1, (m, k)).astype(np.float32))
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    t = qmm.QTensor(gpu_ctx, qtype, k, m)
    try:
        t.set(wire)
        gpu_ctx.set_option("gemm_exact", 1)
        exact = gpu_ctx.mul_mat(t, x, flags=qmm.MM_FORCE_GEMM)
        gpu_ctx.set_option("gemm_exact", 0)
        got = gpu_ctx.mul_mat(t, x)                      # the default dispatch
        assert got.shape == exact.shape and np.isfinite(got).all()
        assert nmse(got, exact) <= F16_GEMM_NMSE, f"fp16 path vs exact int8 path: nmse {nmse(got, exact)}"
        again = gpu_ctx.mul_mat(t, x)
        assert np.array_equal(got, again), "the split-k reduction must be deterministic"
        rows = np.arange(m) if not big else np.unique(np.concatenate([[0, m - 1], rng.integers(0, m, 64)]))
        ref = oracle.mul_mat(qtype, np.ascontiguousarray(wire[rows]), k, len(rows), 1, 1, x[None, None])[0, 0]
        err = nmse(got[:, rows], ref)
        assert err <= MUL_MAT_NMSE_TOL and err <= F16_GEMM_NMSE, err
        # every element, not just the norm: the worst element error relative to the output scale
        scale = np.sqrt(np.mean(ref.astype(np.float64) ** 2))
        assert np.max(np.abs(got[:, rows].astype(np.float64) - ref)) <= 2e-2 * scale
    finally:
        gpu_ctx.set_option("gemm_exact", 0)
        t.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_f16_gemm_batch_broadcast_and_strided_src1(gpu_ctx, qmm, oracle, qtype):
    """src0 [k, m, 2, 1] broadcast over src1 [k, n, 4, 2] (r2 = 2, r3 = 2) with padded src1 rows (nb11 > k * 4)."""
    m, k, n, stride = 192, 320, 40, 320 + 32
    rng = np.random.default_rng(5)
    wire = oracle.quantize_weights(qtype, rng.uniform(-1, 1, (2 * m, k)).astype(np.float32))
    t = qmm.QTensor(gpu_ctx, qtype, k, m, 2, 1)
    t.set(wire)
    xpad = rng.uniform(-1, 1, (2, 4, n, stride)).astype(np.float32)
    xd = gpu_ctx.to_device(xpad)
    out = gpu_ctx.alloc(2 * 4 * n * m * 4)
    try:
        gpu_ctx.mul_mat_device(t, xd.ptr, n, out.ptr, ne12=4, ne13=2, nb11=stride * 4, nb12=n * stride * 4, nb13=4 * n * stride * 4)
        gpu_ctx.synchronize()
        got = out.download(np.float32, 2 * 4 * n * m).reshape(2, 4, n, m)
        ref = oracle.mul_mat(qtype, wire, k, m, 2, 1, np.ascontiguousarray(xpad[..., :k]))
        assert nmse(got, ref) <= F16_GEMM_NMSE
    finally:
        xd.free(); out.free(); t.free()


def test_f16_gemm_extreme_scales(gpu_ctx, qmm, oracle):
    """activation rows spanning 1e-3 .. 1e3 and a zero row: per-block scales keep fp16 in range; zeros stay exactly zero"""
    m, k, n = 384, 512, 64
    rng = np.random.default_rng(11)
    wire = oracle.quantize_weights(Q4_0, rng.uniform(-1, 1, (m, k)).astype(np.float32))
    t = qmm.QTensor(gpu_ctx, Q4_0, k, m)
    t.set(wire)
    x = (rng.standard_normal((n, k)) * np.logspace(-3, 3, n)[:, None]).astype(np.float32)
    x[7] = 0.0
    try:
        got = gpu_ctx.mul_mat(t, x)
        ref = oracle.mul_mat(Q4_0, wire, k, m, 1, 1, x[None, None])[0, 0]
        assert not got[7].any()
        for c in range(n):
            if c != 7:
                assert nmse(got[c], ref[c]) <= F16_GEMM_NMSE, (c, nmse(got[c], ref[c]))
    finally:
        t.free()
