"""Parity of the CUDA path against the oracle, through the C ABI (include/ggml_b200.h), on a real B200.
Bars (north_star): quantize_row_q8_0 bit-exact; per-block int32 dots bit-exact; fp32 mul_mat output within
test-backend-ops' NMSE <= 5e-4 (tests/test-backend-ops.cpp:921-923)."""
import numpy as np
import pytest

from conftest import Q4_0, Q8_0, WIRE, MUL_MAT_NMSE_TOL, F16_GEMM_NMSE, nmse

pytestmark = pytest.mark.gpu
NAMES = {Q4_0: "q4_0", Q8_0: "q8_0"}


def make_w(oracle, qmm, ctx, qtype, m, k, seed, ne02=1, ne03=1):
    rng = np.random.default_rng(seed)
    wire = oracle.quantize_weights(qtype, rng.uniform(-1, 1, (ne03 * ne02 * m, k)))
    t = qmm.QTensor(ctx, qtype, k, m, ne02, ne03)
    t.set(wire)
    return t, wire


# ---- quantize_row_q8_0 -----------------------------------------------------------------------------

def test_quantize_golden_bit_exact(gpu_ctx, golden):
    got = gpu_ctx.quantize_row_q8_0(golden["act_x"])
    assert np.array_equal(got, golden["act_q8_0"])


@pytest.mark.parametrize("k,nrows", [(32, 1), (64, 3), (4096, 17), (768, 128), (16384, 5)])
def test_quantize_random_bit_exact(gpu_ctx, oracle, k, nrows):
    rng = np.random.default_rng(k + nrows)
    x = (rng.standard_normal((nrows, k)) * rng.uniform(1e-4, 1e4, (nrows, 1))).astype(np.float32)
    x[0, :32] = 0
    ref = oracle.quantize_row_q8_0(x)
    assert np.array_equal(gpu_ctx.quantize_row_q8_0(x), ref)
    qs, d = gpu_ctx.quantize_q8_0_planar(x)
    refb = ref.reshape(nrows, k // 32, 34)
    assert np.array_equal(d, refb[:, :, :2].copy().view(np.uint16).reshape(nrows, -1))
    assert np.array_equal(qs.view(np.uint8).reshape(nrows, k // 32, 32), refb[:, :, 2:])


def test_quantize_many_elements_bit_exact(gpu_ctx, oracle):
    """3.1 M elements: the scale at which the scalar _reference quantizer is known to differ."""
    rng = np.random.default_rng(2024)
    x = rng.uniform(-1, 1, (768, 4096)).astype(np.float32)
    assert np.array_equal(gpu_ctx.quantize_row_q8_0(x), oracle.quantize_row_q8_0(x))


# ---- repack: set_tensor / get_tensor round trip ------------------------------------------------------

@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("m,k", [(1, 32), (16, 256), (257, 96), (4096, 4096)])
def test_set_get_roundtrip(gpu_ctx, qmm, qtype, m, k):
    rng = np.random.default_rng(m * 31 + k)
    wire = rng.integers(0, 256, size=(m, k // 32 * WIRE[qtype]), dtype=np.uint8)
    t = qmm.QTensor(gpu_ctx, qtype, k, m)
    t.set(wire)
    assert np.array_equal(t.get(), wire.ravel())
    t.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_set_get_partial_ranges(gpu_ctx, qmm, qtype):
    rng = np.random.default_rng(9)
    m, k = 64, 256
    nb = k // 32
    wb = WIRE[qtype]
    wire = rng.integers(0, 256, size=(m * nb, wb), dtype=np.uint8)
    t = qmm.QTensor(gpu_ctx, qtype, k, m)
    t.set(np.zeros_like(wire))
    t.set(wire[100:300], block_off=100)      # a block-aligned sub-range (ggml set_tensor with offset)
    got = t.get().reshape(-1, wb)
    assert np.array_equal(got[100:300], wire[100:300]) and not got[:100].any() and not got[300:].any()
    assert np.array_equal(t.get(block_off=150, nblocks=7).reshape(-1, wb), wire[150:157])
    t.free()


def test_repacked_layout_is_planes(gpu_ctx, qmm):
    """Device bytes really are [qs plane | fp16 d plane] (what the kernels and TMA descriptors assume)."""
    rng = np.random.default_rng(3)
    m, k = 8, 128
    for qtype, qsb in ((Q4_0, 16), (Q8_0, 32)):
        wb = WIRE[qtype]
        wire = rng.integers(0, 256, size=(m * k // 32, wb), dtype=np.uint8)
        t = qmm.QTensor(gpu_ctx, qtype, k, m)
        t.set(wire)
        raw = t.buf.download(np.uint8, t.nbytes)
        nblk = m * k // 32
        assert np.array_equal(raw[: nblk * qsb].reshape(nblk, qsb), wire[:, 2:])
        assert np.array_equal(raw[nblk * qsb:].reshape(nblk, 2), wire[:, :2])
        t.free()


# ---- per-block int32 dots ---------------------------------------------------------------------------

@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("m,k,n", [(16, 256, 1), (33, 96, 5), (64, 4096, 3), (5, 32, 8), (40, 1024, 11), (300, 768, 1),
                                   (9, 16384, 1), (20, 8192, 2), (7, 12288, 1)])
def test_block_dots_gemv_bit_exact(gpu_ctx, qmm, oracle, qtype, m, k, n):
    t, wire = make_w(oracle, qmm, gpu_ctx, qtype, m, k, seed=m + k + n)
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    ref = oracle.block_dots(qtype, wire, oracle.quantize_row_q8_0(x), k)
    try:
        for stream in (1, 0):
            gpu_ctx.set_option("gemv_stream", stream)
            assert np.array_equal(gpu_ctx.block_dots(t, x, path=0), ref), stream
    finally:
        gpu_ctx.set_option("gemv_stream", 1)
    t.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("m,k,n", [(128, 128, 128), (16, 256, 16), (130, 96, 70), (300, 4096, 33), (257, 32, 129), (64, 1024, 200)])
def test_block_dots_gemm_bit_exact(gpu_ctx, qmm, oracle, qtype, m, k, n):
    """The int32 accumulators the tcgen05 MMAs leave in TMEM, one K=32 MMA per quant block, read back unscaled."""
    t, wire = make_w(oracle, qmm, gpu_ctx, qtype, m, k, seed=m + k + n + 1)
    rng = np.random.default_rng(n + 5)
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    ref = oracle.block_dots(qtype, wire, oracle.quantize_row_q8_0(x), k)
    got = gpu_ctx.block_dots(t, x, path=1)
    assert np.array_equal(got, ref)
    t.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("m,k,n", [(128, 128, 128), (1, 32, 9), (16, 256, 16), (300, 4096, 33), (11008 // 8, 4096, 512), (4096, 1024, 64),
                                   (257, 736, 129), (50257 // 16, 768, 128)])
def test_mul_mat_gemm_vs_oracle(gpu_ctx, qmm, oracle, qtype, m, k, n):
    """Prefill shapes through the tensor-core GEMMs (forced, so small n is covered too): the exact kernel (int8 MMA per quant
    block, fp32 scaling: only the summation order differs from the oracle) and the default, the fp16 contraction (operands rounded
    once to fp16: NMSE ~1e-7, bound 1e-6 here against the reference's 5e-4) at every n since profiles/r02_sweep_n.log."""
    t, wire = make_w(oracle, qmm, gpu_ctx, qtype, m, k, seed=m * 3 + k + n)
    rng = np.random.default_rng(m + n + 2)
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    ref = oracle.mul_mat(qtype, wire, k, m, 1, 1, x[None, None])[0, 0]
    try:
        gpu_ctx.set_option("gemm_exact", 1)
        got = gpu_ctx.mul_mat(t, x, flags=qmm.MM_FORCE_GEMM)
        err = nmse(got, ref)
        assert np.all(np.isfinite(got)) and err <= MUL_MAT_NMSE_TOL and err <= 1e-9, err
        gpu_ctx.set_option("gemm_exact", 0)
        got16 = gpu_ctx.mul_mat(t, x, flags=qmm.MM_FORCE_GEMM)
        err16 = nmse(got16, ref)
        assert np.all(np.isfinite(got16)) and err16 <= MUL_MAT_NMSE_TOL and err16 <= F16_GEMM_NMSE, err16
    finally:
        gpu_ctx.set_option("gemm_exact", 0)
        t.free()


# ---- mul_mat ----------------------------------------------------------------------------------------

def test_mul_mat_golden_cases(gpu_ctx, qmm, golden):
    """The reference CPU backend's own outputs for the test-backend-ops shapes (incl. batch + broadcast)."""
    for ci, (t, m, n, k, bs0, bs1, nr0, nr1) in enumerate(golden["mm_cases"].tolist()):
        w = qmm.QTensor(gpu_ctx, t, k, m, bs0, bs1)
        w.set(golden[f"mm{ci}_a"])
        b = golden[f"mm{ci}_b"].astype(np.float32)
        got = gpu_ctx.mul_mat(w, b)
        ref = golden[f"mm{ci}_out"]
        assert got.shape == ref.shape
        assert np.all(np.isfinite(got))
        err = nmse(got, ref)
        assert err <= MUL_MAT_NMSE_TOL, f"case {ci} {(t, m, n, k, bs0, bs1, nr0, nr1)}: nmse {err}"
        assert err <= (1e-9 if n <= 8 and n * k < 32768 else F16_GEMM_NMSE), f"case {ci}: only fp32 summation order (GEMV) / one fp16 rounding per operand (tensor-core path) may differ, nmse {err}"
        w.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("m,k,n", [(1, 32, 1), (16, 256, 1), (16, 256, 16), (50257 // 64, 768, 1), (4096, 4096, 1),
                                   (1000, 3072, 2), (333, 768, 7), (129, 64, 8), (64, 16384, 1), (2304, 768, 9), (4096, 16384, 1),
                                   (100, 8192, 3), (50, 12288, 1), (700, 28672 // 7 * 8, 1), (148, 256, 1), (147, 512, 8)])
def test_mul_mat_vs_oracle(gpu_ctx, qmm, oracle, qtype, m, k, n):
    t, wire = make_w(oracle, qmm, gpu_ctx, qtype, m, k, seed=m * 7 + k + n)
    rng = np.random.default_rng(m + n)
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    ref = oracle.mul_mat(qtype, wire, k, m, 1, 1, x[None, None])[0, 0]
    try:
        for stream in (1, 0):                       # streaming GEMV (b200_gemv_stream.cu) and generic GEMV (b200_gemv.cu)
            gpu_ctx.set_option("gemv_stream", stream)
            for flags in (0, qmm.MM_FORCE_GEMV):
                got = gpu_ctx.mul_mat(t, x, flags=flags)
                err = nmse(got, ref)
                # the default dispatch leaves the GEMV (fp32 summation order only) for the fp16 tensor-core path past 8 columns
                bound = 1e-9 if flags == qmm.MM_FORCE_GEMV or (n <= 8 and n * k < 32768) else F16_GEMM_NMSE
                assert err <= MUL_MAT_NMSE_TOL and err <= bound, (stream, flags, err)
    finally:
        gpu_ctx.set_option("gemv_stream", 1)
    assert nmse(gpu_ctx.mul_mat_host(t, x), ref) <= (1e-9 if n <= 8 and n * k < 32768 else F16_GEMM_NMSE)
    t.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_mul_mat_strided_src1(gpu_ctx, qmm, oracle, qtype):
    """src1 rows may be strided (nb11 > k*4), src/ggml.c:11956-11971."""
    m, k, n, stride = 48, 256, 4, 256 + 64
    t, wire = make_w(oracle, qmm, gpu_ctx, qtype, m, k, seed=11)
    rng = np.random.default_rng(12)
    xs = rng.uniform(-1, 1, (n, stride)).astype(np.float32)
    xd = gpu_ctx.to_device(xs)
    out = gpu_ctx.alloc(n * m * 4)
    gpu_ctx.mul_mat_device(t, xd.ptr, n, out.ptr, nb11=stride * 4)
    gpu_ctx.synchronize()
    got = out.download(np.float32, n * m).reshape(n, m)
    ref = oracle.mul_mat(qtype, wire, k, m, 1, 1, np.ascontiguousarray(xs[:, :k])[None, None])[0, 0]
    assert nmse(got, ref) <= 1e-9
    for b in (t, xd, out):
        b.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_mul_mat_row_range_view(gpu_ctx, qmm, oracle, qtype):
    """src0 given as a row-range of a bigger repacked tensor (what the row-split across GPUs uses)."""
    m, k, n = 96, 512, 2
    t, wire = make_w(oracle, qmm, gpu_ctx, qtype, m, k, seed=21)
    rng = np.random.default_rng(22)
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    ref = oracle.mul_mat(qtype, wire, k, m, 1, 1, x[None, None])[0, 0]
    xd = gpu_ctx.to_device(x)
    r0, r1 = 32, 80
    out = gpu_ctx.alloc(n * (r1 - r0) * 4)
    gpu_ctx.mul_mat_device(t, xd.ptr, n, out.ptr, block_off=r0 * (k // 32), m=r1 - r0)
    gpu_ctx.synchronize()
    got = out.download(np.float32, n * (r1 - r0)).reshape(n, r1 - r0)
    assert nmse(got, ref[:, r0:r1]) <= 1e-9
    for b in (t, xd, out):
        b.free()


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("k,ms", [(4096, [4096, 4096, 4096, 16384]), (768, [2304, 3072]), (1024, [5, 300, 1]), (256, [148, 149, 147, 2000])])
def test_mul_mat_batch_same_input(gpu_ctx, qmm, oracle, qtype, k, ms):
    """b200_mul_mat_batch: independent decode mul_mats on the same src1 in one launch == the separate calls, bit for bit."""
    rng = np.random.default_rng(k + len(ms))
    x = rng.uniform(-1, 1, (1, k)).astype(np.float32)
    xd = gpu_ctx.to_device(x)
    ws, outs, refs = [], [], []
    for i, m in enumerate(ms):
        t, wire = make_w(oracle, qmm, gpu_ctx, qtype, m, k, seed=100 + i)
        ws.append(t)
        outs.append(gpu_ctx.alloc(m * 4))
        refs.append(gpu_ctx.mul_mat(t, x)[0])
    gpu_ctx.mul_mat_batch([gpu_ctx.make_args(t, xd.ptr, 1, o.ptr) for t, o in zip(ws, outs)])
    gpu_ctx.synchronize()
    for m, o, ref in zip(ms, outs, refs):
        got = o.download(np.float32, m)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # a mixed batch (different k / n > 1 entries) silently falls back to one launch per entry
    t2, wire2 = make_w(oracle, qmm, gpu_ctx, qtype, 64, 512, seed=7)
    x2 = rng.uniform(-1, 1, (3, 512)).astype(np.float32)
    x2d = gpu_ctx.to_device(x2)
    o2 = gpu_ctx.alloc(3 * 64 * 4)
    gpu_ctx.mul_mat_batch([gpu_ctx.make_args(ws[0], xd.ptr, 1, outs[0].ptr), gpu_ctx.make_args(t2, x2d.ptr, 3, o2.ptr)])
    gpu_ctx.synchronize()
    ref2 = oracle.mul_mat(qtype, wire2, 512, 64, 1, 1, x2[None, None])[0, 0]
    assert nmse(o2.download(np.float32, 3 * 64).reshape(3, 64), ref2) <= 1e-9
    for b in ws + outs + [xd, t2, x2d, o2]:
        b.free()


def test_mul_mat_errors(gpu_ctx, qmm):
    """Same rejections as the reference's asserts; reported as error codes, never a silent fallback."""
    t = qmm.QTensor(gpu_ctx, Q4_0, 64, 4)
    x = gpu_ctx.alloc(1024)
    with pytest.raises(qmm.B200Error) as e:
        a = qmm.MulMatArgs()
        a.type, a.src0_dev, a.src0_nblocks_total = 1, t.ptr, t.nblocks   # F16 weights: outside the path
        a.ne00, a.ne01, a.ne02, a.ne03, a.ne11, a.ne12, a.ne13 = 64, 4, 1, 1, 1, 1, 1
        a.src1_dev, a.dst_dev, a.nb11, a.nb12, a.nb13 = x.ptr, x.ptr, 256, 256, 256
        gpu_ctx._check(gpu_ctx.lib.b200_mul_mat(gpu_ctx.h, a))
    assert e.value.code == qmm.ERR_UNSUPPORTED
    with pytest.raises(qmm.B200Error) as e:
        t2 = qmm.QTensor(gpu_ctx, Q4_0, 64, 4, ne02=2)
        gpu_ctx.mul_mat_device(t2, x.ptr, 1, x.ptr, ne12=3)      # 3 % 2 != 0 -> ggml_can_mul_mat false
    assert e.value.code == qmm.ERR_INVALID
    gpu_ctx.synchronize()


# ---- size-independent properties at BASELINE.json's full sizes ------------------------------------------

@pytest.mark.parametrize("qtype,m,k,n", [(Q4_0, 4096, 4096, 1), (Q4_0, 11008, 4096, 512), (Q8_0, 11008, 4096, 512),
                                         (Q4_0, 50400, 4096, 1), (Q4_0, 28672, 8192, 1)])
def test_full_size_properties(gpu_ctx, qmm, oracle, qtype, m, k, n):
    """(1) power-of-two scaling of x scales dst exactly (Q8_0 quantization is scale-equivariant for 2^e);
    (2) zero activations give exactly zero; (3) a sample of rows/columns matches the oracle;
    (4) get_tensor returns the exact wire bytes (checksum)."""
    wire = qmm.random_wire_weights(qtype, k, m, seed=m + k)
    t = qmm.QTensor(gpu_ctx, qtype, k, m)
    t.set(wire)
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    y1 = gpu_ctx.mul_mat(t, x)
    assert np.all(np.isfinite(y1))
    y4 = gpu_ctx.mul_mat(t, x * np.float32(4.0))
    assert np.array_equal(y4, y1 * np.float32(4.0))
    assert not gpu_ctx.mul_mat(t, np.zeros_like(x)).any()
    rows = np.unique(np.concatenate([[0, m - 1], rng.integers(0, m, 48)]))
    cols = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, 6)]))
    ref = oracle.mul_mat(qtype, np.ascontiguousarray(wire[rows]), k, len(rows), 1, 1, np.ascontiguousarray(x[cols])[None, None])[0, 0]
    sub = y1[np.ix_(cols, rows)]
    assert nmse(sub, ref) <= MUL_MAT_NMSE_TOL and nmse(sub, ref) <= (1e-8 if n <= 8 and n * k < 32768 else F16_GEMM_NMSE)
    back = t.get()
    assert int(back.astype(np.uint64).sum()) == int(wire.astype(np.uint64).sum()) and np.array_equal(back[:4096], wire.ravel()[:4096])
    t.free()
