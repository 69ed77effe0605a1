"""The operators either side of the quantized mul_mat, through the C ABI (b200_op_*), against numpy restatements of
the reference CPU semantics (src/ggml.c, cited per test) on the same inputs.  Float work: tolerance 1e-6 NMSE unless the reference itself
rounds through fp16 tables (GELU, SOFT_MAX: src/ggml.c:1978-1991, :13470-13483), where the bar is test-backend-ops' own (1e-6 for
SOFT_MAX, tests/test-backend-ops.cpp:1097-1099).  The reference's own harness runs the same ops in tests/test_gpu_backend_ops.py."""
import numpy as np
import pytest

from conftest import Q4_0, Q8_0, nmse
from rope_restatement import rope_numpy

pytestmark = pytest.mark.gpu
TOL = 1e-6


def up(qmm, ctx, a):
    return qmm.DTensor.from_numpy(ctx, np.ascontiguousarray(a))


def empty(qmm, ctx, shape, ttype=0):
    return qmm.DTensor(ctx, ttype, list(shape[::-1]))


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("batch", [1, 3])
def test_get_rows_from_repacked_quantized_rows(qmm, gpu_ctx, oracle, qtype, batch):
    """ggml_compute_forward_get_rows_q (src/ggml.c:12874): dequantize_row of the selected rows, bit for bit"""
    rng = np.random.default_rng(qtype + batch)
    k, m, r = 768, 500, 37
    w = rng.uniform(-1, 1, (batch, m, k)).astype(np.float32)
    wire = oracle.quantize_weights(qtype, w)
    qt = qmm.QTensor(gpu_ctx, qtype, k, m, batch)
    qt.set(wire)
    rows = rng.integers(0, m, (batch, r)).astype(np.int32)
    dst = empty(qmm, gpu_ctx, (batch, r, k))
    gpu_ctx.op_get_rows(qt, up(qmm, gpu_ctx, rows), dst)
    gpu_ctx.synchronize()
    deq = oracle.dequantize(qtype, wire, k).reshape(batch, m, k)
    want = np.stack([deq[b][rows[b]] for b in range(batch)])
    assert np.array_equal(dst.numpy().reshape(batch, r, k), want)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("k", [768, 10, 1])
def test_get_rows_dense_sources(qmm, gpu_ctx, dtype, k):
    rng = np.random.default_rng(k)
    a = rng.uniform(-1, 1, (2, 50, k)).astype(dtype)
    rows = rng.integers(0, 50, (2, 7)).astype(np.int32)
    dst = empty(qmm, gpu_ctx, (2, 7, k))
    gpu_ctx.op_get_rows(up(qmm, gpu_ctx, a), up(qmm, gpu_ctx, rows), dst)
    gpu_ctx.synchronize()
    want = np.stack([a[b][rows[b]] for b in range(2)]).astype(np.float32)
    assert np.array_equal(dst.numpy().reshape(2, 7, k), want)


@pytest.mark.parametrize("ne,nr", [((16, 10, 10, 10), (1, 1, 1, 1)), ((16, 10, 10, 10), (2, 1, 1, 2)), ((1, 1, 8, 1), (5, 3, 1, 1)), ((768, 1, 1, 1), (1, 128, 1, 1)),
                                   ((2304, 1, 1, 1), (1, 1, 1, 1)), ((3, 5, 1, 1), (1, 2, 3, 1))])
@pytest.mark.parametrize("op", ["add", "mul", "div"])
def test_binary_broadcast(qmm, gpu_ctx, op, ne, nr):
    """ggml_compute_forward_add_f32 (src/ggml.c:8568): src1 repeated over src0 in every dimension; exact in fp32"""
    rng = np.random.default_rng(sum(ne) + sum(nr))
    big = tuple(n * r for n, r in zip(ne, nr))[::-1]
    a = rng.uniform(1, 2, big).astype(np.float32)
    b = rng.uniform(1, 2, ne[::-1]).astype(np.float32)
    dst = empty(qmm, gpu_ctx, big)
    gpu_ctx.op_binary({"add": qmm.OP_ADD, "mul": qmm.OP_MUL, "div": qmm.OP_DIV}[op], up(qmm, gpu_ctx, a), up(qmm, gpu_ctx, b), dst)
    gpu_ctx.synchronize()
    bb = np.tile(b, nr[::-1])
    want = {"add": a + bb, "mul": a * bb, "div": a / bb}[op]
    assert np.array_equal(dst.numpy(), want)


def test_binary_in_place_and_strided(qmm, gpu_ctx):
    """dst == src0 (ggml_gallocr computes ADD in place) and a src0 that is a column slice of a wider matrix"""
    rng = np.random.default_rng(3)
    wide = rng.uniform(-1, 1, (12, 2304)).astype(np.float32)
    t = up(qmm, gpu_ctx, wide)
    v = t.view([768, 12, 1, 1], [4, 2304 * 4, 2304 * 4 * 12, 2304 * 4 * 12], offset=768 * 4)      # the K slice of a fused qkv result
    bias = rng.uniform(-1, 1, (768,)).astype(np.float32)
    gpu_ctx.op_binary(qmm.OP_ADD, v, up(qmm, gpu_ctx, bias), v)
    gpu_ctx.synchronize()
    want = wide.copy()
    want[:, 768:1536] += bias
    assert np.array_equal(t.numpy().reshape(12, 2304), want)


GELU_C = 0.79788456080286535587989211986876


@pytest.mark.parametrize("name", ["gelu", "gelu_quick", "silu", "relu", "tanh", "sigmoid", "neg", "abs", "step", "sgn", "elu", "hardswish", "hardsigmoid"])
def test_unary(qmm, gpu_ctx, name):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-150, 150, 4000), rng.uniform(-3, 3, 4000), [0.0, -0.0, 10.0, -10.0]]).astype(np.float32).reshape(4, -1)
    dst = empty(qmm, gpu_ctx, x.shape)
    gpu_ctx.op_unary(name, up(qmm, gpu_ctx, x), dst)
    gpu_ctx.synchronize()
    x64 = x.astype(np.float64)
    with np.errstate(over="ignore"):
        want = {"gelu": 0.5 * x64 * (1 + np.tanh(GELU_C * x64 * (1 + 0.044715 * x64 * x64))), "gelu_quick": x64 / (1 + np.exp(1.702 * -x64)),
                "silu": x64 / (1 + np.exp(-x64)), "relu": np.maximum(x64, 0), "tanh": np.tanh(x64), "sigmoid": 1 / (1 + np.exp(-x64)), "neg": -x64,
                "abs": np.abs(x64), "step": (x64 > 0).astype(np.float64), "sgn": np.sign(x64), "elu": np.where(x64 > 0, x64, np.expm1(x64)),
                "hardswish": x64 * np.clip((x64 + 3) / 6, 0, 1), "hardsigmoid": np.clip((x64 + 3) / 6, 0, 1)}[name]
    got = dst.numpy()
    assert np.isfinite(got).all()
    assert nmse(got, want) <= 1e-10, nmse(got, want)


@pytest.mark.parametrize("ne0", [64, 768, 1000, 1028, 2048, 4096, 5000, 8192, 8196])       # warp kernel <= 1024, row-in-registers CTA <= 8192, loop kernel beyond / unaligned
@pytest.mark.parametrize("rms", [False, True])
def test_norm_and_fused_affine(qmm, gpu_ctx, ne0, rms):
    """ggml_compute_forward_norm_f32 (src/ggml.c:11353) / rms_norm; and norm * gain + bias in one kernel"""
    rng = np.random.default_rng(ne0)
    x = (rng.uniform(-1, 1, (3, 5, ne0)) + 0.3).astype(np.float32)
    g = rng.uniform(0.9, 1.1, (ne0,)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (ne0,)).astype(np.float32)
    eps = 1e-5
    x64 = x.astype(np.float64)
    if rms:
        want = x64 / np.sqrt((x64 * x64).mean(-1, keepdims=True) + eps)
    else:
        c = x64 - x64.mean(-1, keepdims=True)
        want = c / np.sqrt((c * c).mean(-1, keepdims=True) + eps)
    tx = up(qmm, gpu_ctx, x)
    dst = empty(qmm, gpu_ctx, x.shape)
    gpu_ctx.op_norm(tx, dst, eps, rms=rms)
    gpu_ctx.synchronize()
    assert nmse(dst.numpy(), want) <= 1e-12
    gpu_ctx.op_norm(tx, tx, eps, gain=up(qmm, gpu_ctx, g), bias=up(qmm, gpu_ctx, b), rms=rms)          # in place
    gpu_ctx.synchronize()
    assert nmse(tx.numpy(), want * g + b) <= 1e-12


def test_scale_and_diag_mask_inf(qmm, gpu_ctx):
    """ggml_compute_forward_scale_f32 (src/ggml.c:12637), ggml_compute_forward_diag_mask_f32 (:13301)"""
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, (2, 12, 10, 23)).astype(np.float32)
    tx = up(qmm, gpu_ctx, x)
    dst = empty(qmm, gpu_ctx, x.shape)
    gpu_ctx.op_scale(tx, dst, 0.125)
    gpu_ctx.synchronize()
    assert np.array_equal(dst.numpy(), x * np.float32(0.125))
    n_past = 13
    gpu_ctx.op_diag_mask_inf(tx, tx, n_past)
    gpu_ctx.synchronize()
    want = x.copy()
    i = np.arange(23)[None, :]
    j = np.arange(10)[:, None]
    want[:, :, i > n_past + j] = -np.inf
    assert np.array_equal(tx.numpy(), want)


def ref_soft_max(x, mask, scale, max_bias):
    """ggml_compute_forward_soft_max_f32 (src/ggml.c:13393) in float64, without its fp16 exp table"""
    ne3, ne2, ne1, ne0 = x.shape
    n_head_log2 = 1 << int(np.floor(np.log2(ne2)))
    m0, m1 = 2.0 ** (-max_bias / n_head_log2), 2.0 ** (-(max_bias / 2.0) / n_head_log2)
    w = x.astype(np.float64) * scale
    if mask is not None:
        for h in range(ne2):
            slope = 1.0 if max_bias <= 0 else (m0 ** (h + 1) if h < n_head_log2 else m1 ** (2 * (h - n_head_log2) + 1))
            w[:, h] += slope * mask.astype(np.float64)[:ne1]
    w -= w.max(-1, keepdims=True)
    e = np.exp(w)
    return e / e.sum(-1, keepdims=True)


@pytest.mark.parametrize("ne0,ne1", [(16, 16), (1023, 15), (1024, 1024), (5000, 3), (1, 4)])
@pytest.mark.parametrize("mask,max_bias", [(False, 0.0), (True, 0.0), (True, 8.0)])
def test_soft_max(qmm, gpu_ctx, ne0, ne1, mask, max_bias):
    rng = np.random.default_rng(ne0 + ne1)
    heads = 1 if ne0 * ne1 > 100000 else 6
    x = rng.uniform(-1, 1, (1, heads, ne1, ne0)).astype(np.float32)
    mk = rng.uniform(-1, 1, (ne1, ne0)).astype(np.float32) if mask else None
    tx = up(qmm, gpu_ctx, x)
    dst = empty(qmm, gpu_ctx, x.shape)
    gpu_ctx.op_soft_max(tx, dst, mask=up(qmm, gpu_ctx, mk) if mask else None, scale=0.1, max_bias=max_bias)
    gpu_ctx.synchronize()
    want = ref_soft_max(x, mk, 0.1, max_bias)
    assert nmse(dst.numpy(), want) <= 1e-10
    if mask:
        gpu_ctx.op_soft_max(tx, dst, mask=up(qmm, gpu_ctx, mk.astype(np.float16)), scale=0.1, max_bias=max_bias)
        gpu_ctx.synchronize()
        assert nmse(dst.numpy(), ref_soft_max(x, mk.astype(np.float16), 0.1, max_bias)) <= 1e-10


@pytest.mark.parametrize("n_past,N", [(0, 128), (128, 1), (5, 3)])
def test_fused_scale_mask_soft_max_equals_the_three_ops(qmm, gpu_ctx, n_past, N):
    """SCALE -> DIAG_MASK_INF -> SOFT_MAX (examples/gpt-2/main-backend.cpp:567-583) as one kernel, in place: bitwise what the three
    kernels give one after the other"""
    rng = np.random.default_rng(n_past + N)
    x = rng.uniform(-4, 4, (1, 12, N, n_past + N)).astype(np.float32)
    s = 1.0 / np.sqrt(64.0)
    a = up(qmm, gpu_ctx, x)
    gpu_ctx.op_scale(a, a, s)
    gpu_ctx.op_diag_mask_inf(a, a, n_past)
    gpu_ctx.op_soft_max(a, a)
    b = up(qmm, gpu_ctx, x)
    gpu_ctx.op_soft_max(b, b, scale=s, n_past=n_past)
    gpu_ctx.synchronize()
    ga, gb = a.numpy(), b.numpy()
    assert np.array_equal(ga, gb)
    j = np.arange(N)[:, None]
    i = np.arange(n_past + N)[None, :]
    assert (gb[..., i > n_past + j] == 0).all() and np.allclose(gb.sum(-1), 1, atol=1e-5)


@pytest.mark.parametrize("src,dst", [(np.float32, np.float32), (np.float32, np.float16), (np.float16, np.float32), (np.float16, np.float16), (np.int32, np.int32),
                                     (np.int16, np.int16)])
def test_copy_permuted_to_contiguous_and_into_a_view(qmm, gpu_ctx, src, dst):
    """ggml_compute_forward_dup (src/ggml.c:8535): CONT of a permuted view (the V cache transposed, main-backend.cpp:588-595), and CPY of a
    strided 2-D slice into a 1-D view of the KV cache (:531-535)"""
    rng = np.random.default_rng(1)
    T, H, D = 9, 12, 64
    a = (rng.uniform(-100, 100, (T, H, D))).astype(src)
    ta = up(qmm, gpu_ctx, a)
    es = a.itemsize
    # permute(1, 2, 0, 3) of [D, H, T]: new axes (ne0 = T, ne1 = D, ne2 = H)
    pv = ta.view([T, D, H, 1], [es * D * H, es, es * D, es * D * H * T])
    ttype = {np.dtype(np.float32): qmm.TYPE_F32, np.dtype(np.float16): qmm.TYPE_F16, np.dtype(np.int32): qmm.TYPE_I32, np.dtype(np.int16): qmm.TYPE_I16}[np.dtype(dst)]
    out = qmm.DTensor(gpu_ctx, ttype, [T, D, H, 1])
    gpu_ctx.op_copy(pv, out)
    gpu_ctx.synchronize()
    want = np.transpose(a, (1, 2, 0)).astype(dst)           # [H][D][T]
    assert np.array_equal(out.numpy().reshape(H, D, T), want)
    # a column slice [D*H, T] with row stride 3*D*H  ->  a flat run inside a bigger buffer at an offset
    wide = rng.uniform(-100, 100, (T, 3 * H * D)).astype(src)
    tw = up(qmm, gpu_ctx, wide)
    sl = tw.view([H * D, T, 1, 1], [es, es * 3 * H * D, es * 3 * H * D * T, es * 3 * H * D * T], offset=es * H * D)
    cache = qmm.DTensor(gpu_ctx, ttype, [4 * T * H * D])
    cache.buf.upload(np.zeros(4 * T * H * D, dst))
    ds = np.dtype(dst).itemsize
    flat = cache.view([T * H * D, 1, 1, 1], [ds, ds * T * H * D, ds * T * H * D, ds * T * H * D], offset=ds * T * H * D)
    gpu_ctx.op_copy(sl, flat)
    gpu_ctx.synchronize()
    got = cache.numpy().reshape(4, T, H * D)
    assert np.array_equal(got[1], wide[:, H * D:2 * H * D].astype(dst)) and not got[0].any() and not got[2:].any()


@pytest.mark.parametrize("atype", [np.float32, np.float16])
@pytest.mark.parametrize("m,n,k,bs,nr", [(16, 1, 256, (10, 10), (2, 2)), (16, 16, 256, (10, 1), (1, 1)), (129, 1, 64, (12, 1), (1, 1)), (128, 128, 64, (12, 1), (1, 1)),
                                         (64, 130, 257, (3, 1), (1, 1)), (257, 7, 33, (1, 1), (1, 1))])
def test_mul_mat_dense(qmm, gpu_ctx, atype, m, n, k, bs, nr):
    """ggml_compute_forward_mul_mat with an F32 / F16 src0 (src/ggml.c:11808), incl. the batch broadcast of test-backend-ops' bs / nr pattern"""
    rng = np.random.default_rng(m + n + k)
    a = rng.uniform(-1, 1, (bs[1], bs[0], m, k)).astype(atype)
    b = rng.uniform(-1, 1, (bs[1] * nr[1], bs[0] * nr[0], n, k)).astype(np.float32)
    dst = empty(qmm, gpu_ctx, (bs[1] * nr[1], bs[0] * nr[0], n, m))
    gpu_ctx.op_mul_mat_dense(up(qmm, gpu_ctx, a), up(qmm, gpu_ctx, b), dst)
    gpu_ctx.synchronize()
    a64 = np.repeat(np.repeat(a.astype(np.float64), nr[1], 0), nr[0], 1)
    want = np.einsum("xymk,xynk->xynm", a64, b.astype(np.float64))
    assert nmse(dst.numpy(), want) <= 1e-10


def test_attention_scores_on_permuted_kv_views(qmm, gpu_ctx):
    """K*Q exactly as gpt-2 builds it (main-backend.cpp:556-567): K = the cache [hd, heads, T] permuted to [hd, T, heads] (no copy), Q likewise"""
    rng = np.random.default_rng(11)
    T, H, D, N = 130, 12, 64, 2
    kc = rng.uniform(-1, 1, (T, H, D)).astype(np.float32)
    qc = rng.uniform(-1, 1, (N, H, D)).astype(np.float32)
    K = up(qmm, gpu_ctx, kc).view([D, T, H, 1], [4, 4 * D * H, 4 * D, 4 * D * H * T])
    Q = up(qmm, gpu_ctx, qc).view([D, N, H, 1], [4, 4 * D * H, 4 * D, 4 * D * H * N])
    dst = empty(qmm, gpu_ctx, (1, H, N, T))
    gpu_ctx.op_mul_mat_dense(K, Q, dst)
    gpu_ctx.synchronize()
    want = np.einsum("thd,nhd->hnt", kc.astype(np.float64), qc.astype(np.float64))
    assert nmse(dst.numpy().reshape(H, N, T), want) <= 1e-10


def test_unsupported_combinations_are_refused(qmm, gpu_ctx):
    a = up(qmm, gpu_ctx, np.zeros((4, 8), np.float32))
    i = up(qmm, gpu_ctx, np.zeros((4, 8), np.int32))
    with pytest.raises(qmm.B200Error) as e:
        gpu_ctx.op_copy(a, i)                           # F32 -> I32
    assert e.value.code == qmm.ERR_UNSUPPORTED
    with pytest.raises(qmm.B200Error):
        gpu_ctx.op_binary(qmm.OP_ADD, a, up(qmm, gpu_ctx, np.zeros((3,), np.float32)), a)      # 8 is not a multiple of 3
    with pytest.raises(qmm.B200Error):
        gpu_ctx.op_norm(a, a, 0.0)                      # the reference asserts eps > 0


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("k,m,n", [(768, 2304, 1), (3072, 768, 1), (768, 3072, 3), (4096, 4096, 1), (96, 50, 2), (16384, 4096, 1)])
@pytest.mark.parametrize("bias,gelu,res", [(True, False, 0), (True, True, 0), (True, False, 1), (False, False, 1), (True, True, 1), (False, False, 2), (True, False, 2)])
def test_mul_mat_fused_epilogue_equals_the_separate_operators(qmm, gpu_ctx, oracle, qtype, k, m, n, bias, gelu, res):
    """b200_mul_mat_fused: dst = act(W x + bias) + residual in the GEMV epilogue (streaming kernel for k % 256 == 0, generic kernel
    otherwise; k > 4096 goes through the k-split combine) -- bitwise what b200_mul_mat, ADD, GELU, ADD give one after the other, also with
    the residual aliasing dst (in place)"""
    rng = np.random.default_rng(k + m + n)
    w = qmm.QTensor(gpu_ctx, qtype, k, m)
    w.set(oracle.quantize_weights(qtype, rng.uniform(-1, 1, (m, k)).astype(np.float32) * (2.0 / np.sqrt(k))))
    x = up(qmm, gpu_ctx, rng.uniform(-1, 1, (n, k)).astype(np.float32))
    b = up(qmm, gpu_ctx, rng.uniform(-1, 1, (m,)).astype(np.float32))
    r0 = rng.uniform(-1, 1, (n, m)).astype(np.float32)
    r2 = up(qmm, gpu_ctx, rng.uniform(-1, 1, (n, m)).astype(np.float32))              # res == 2: the second residual (GPT-J: + MLP branch, + residual stream)
    # separate operators
    d1 = empty(qmm, gpu_ctx, (n, m))
    gpu_ctx.mul_mat_device(w, x.buf.ptr, n, d1.buf.ptr)
    if bias:
        gpu_ctx.op_binary(qmm.OP_ADD, d1, b, d1)
    if gelu:
        gpu_ctx.op_unary("gelu", d1, d1)
    if res:
        gpu_ctx.op_binary(qmm.OP_ADD, d1, up(qmm, gpu_ctx, r0), d1)
    if res == 2:
        gpu_ctx.op_binary(qmm.OP_ADD, d1, r2, d1)
    # one launch; the residual IS the destination buffer
    d2 = up(qmm, gpu_ctx, r0)
    l0 = gpu_ctx.launch_count()
    gpu_ctx.mul_mat_fused(w, x.buf.ptr, n, d2.buf.ptr, bias_ptr=b.buf.ptr if bias else 0, residual_ptr=d2.buf.ptr if res else 0,
                          act=qmm.EPI_GELU if gelu else qmm.EPI_NONE, residual2_ptr=r2.buf.ptr if res == 2 else 0)
    assert gpu_ctx.launch_count() - l0 == 1
    gpu_ctx.synchronize()
    assert np.array_equal(d1.numpy(), d2.numpy())


def test_mul_mat_fused_refuses_prefill_shapes(qmm, gpu_ctx, oracle):
    w = qmm.QTensor(gpu_ctx, Q4_0, 256, 64)
    w.set(oracle.quantize_weights(Q4_0, np.ones((64, 256), np.float32)))
    x = up(qmm, gpu_ctx, np.ones((64, 256), np.float32))
    d = empty(qmm, gpu_ctx, (64, 64))
    with pytest.raises(qmm.B200Error) as e:
        gpu_ctx.mul_mat_fused(w, x.buf.ptr, 64, d.buf.ptr, bias_ptr=d.buf.ptr)
    assert e.value.code == qmm.ERR_UNSUPPORTED


def test_events_order_two_streams(qmm, gpu_ctx):
    """b200_event_*: work recorded on one context's stream, awaited from another context's stream and from the host (the SPI's event entries)"""
    import ctypes as C
    other = qmm.Context(0)
    try:
        n = 1 << 22
        x = np.arange(n, dtype=np.float32)
        a = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [n])
        b = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [n])
        ev = C.c_void_p()
        gpu_ctx._check(gpu_ctx.lib.b200_event_create(gpu_ctx.h, C.byref(ev)))
        for rep in range(5):
            gpu_ctx._check(gpu_ctx.lib.b200_upload_async(gpu_ctx.h, C.c_void_p(a.buf.ptr), x.ctypes.data, x.nbytes))      # stream 1
            gpu_ctx.op_scale(a, a, 2.0)
            gpu_ctx._check(gpu_ctx.lib.b200_event_record(gpu_ctx.h, ev))
            other._check(other.lib.b200_event_wait(other.h, ev))                                                             # stream 2 waits
            other.op_scale(a, b, 0.5)
            other.synchronize()
            assert np.array_equal(b.numpy().reshape(-1), x)
            x += 1.0
        gpu_ctx._check(gpu_ctx.lib.b200_event_record(gpu_ctx.h, ev))
        assert gpu_ctx.lib.b200_event_synchronize(ev) == 0
        gpu_ctx.lib.b200_event_destroy(ev)
    finally:
        other.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("ne0,heads,n_dims,mode,yarn,xpos", [(64, 16, 64, 0, False, False),      # GPT-J: ggml_rope_inplace(.., n_rot = 64, 0, 0), examples/gpt-j/main.cpp:473
                                                             (128, 5, 128, 0, True, False), (64, 7, 64, 2, False, False), (80, 4, 20, 2, True, False),
                                                             (128, 3, 128, 0, False, True)])
def test_rope(qmm, gpu_ctx, dtype, ne0, heads, n_dims, mode, yarn, xpos):
    """GGML_OP_ROPE forward: normal and NeoX pairing, YaRN mixing (ext_factor != 0) and the xPos factor, F32 and F16, also in place"""
    rng = np.random.default_rng(ne0 + heads + mode)
    B, T = 2, 9
    x = rng.uniform(-1, 1, (B, T, heads, ne0)).astype(dtype)
    pos = rng.integers(0, 512, T).astype(np.int32)
    kw = dict(n_dims=n_dims, mode=mode, n_orig_ctx=256 if yarn else 0, freq_base=10000.0, freq_scale=0.5 if yarn else 1.0, ext_factor=0.7 if yarn else 0.0,
              attn_factor=1.1 if yarn else 1.0, beta_fast=32.0 if yarn else 0.0, beta_slow=1.0 if yarn else 0.0)
    if xpos:
        kw.update(xpos_base=512.0, xpos_down=True)
    want = rope_numpy(x, pos, **kw)
    ttype = qmm.TYPE_F32 if dtype == np.float32 else qmm.TYPE_F16
    a, tp = up(qmm, gpu_ctx, x), up(qmm, gpu_ctx, pos)
    dst = qmm.DTensor(gpu_ctx, ttype, [ne0, heads, T, B])
    gpu_ctx.op_rope(a, tp, dst, n_ctx=512, **kw)
    if dtype == np.float32:                                          # the rotation stored as F16 == ROPE then CPY(F32 -> F16), bit for bit
        h1, h2 = qmm.DTensor(gpu_ctx, qmm.TYPE_F16, [ne0, heads, T, B]), qmm.DTensor(gpu_ctx, qmm.TYPE_F16, [ne0, heads, T, B])
        gpu_ctx.op_rope(a, tp, h1, n_ctx=512, **kw)
        gpu_ctx.op_copy(dst, h2)
        gpu_ctx.synchronize()
        assert np.array_equal(h1.numpy(), h2.numpy())
    gpu_ctx.op_rope(a, tp, a, n_ctx=512, **kw)                       # ggml_rope_inplace
    gpu_ctx.synchronize()
    got, got_inplace = dst.numpy().reshape(x.shape), a.numpy().reshape(x.shape)
    assert np.array_equal(got, got_inplace)
    assert nmse(got.astype(np.float64), want.astype(np.float64)) <= (1e-7 if dtype == np.float32 else 1e-6)
    assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= (2e-5 if dtype == np.float32 else 2e-3)


@pytest.mark.parametrize("dtype", [np.float32, np.float16, np.int32, np.int16])
@pytest.mark.parametrize("nr", [(1, 1, 1, 1), (2, 1, 1, 1), (1, 3, 1, 2), (2, 2, 2, 2)])
def test_repeat(qmm, gpu_ctx, dtype, nr):
    """ggml_compute_forward_repeat (src/ggml.c:10323): the tiling np.tile does; GPT-J's old-style bias broadcast (examples/gpt-j/main.cpp:452-456)"""
    rng = np.random.default_rng(7)
    shape = (3, 4, 5, 10)                                             # [ne3][ne2][ne1][ne0]
    a = rng.integers(-1000, 1000, shape).astype(dtype)
    ttype = {np.dtype(np.float32): qmm.TYPE_F32, np.dtype(np.float16): qmm.TYPE_F16, np.dtype(np.int32): qmm.TYPE_I32, np.dtype(np.int16): qmm.TYPE_I16}[np.dtype(dtype)]
    want = np.tile(a, nr[::-1])
    dst = qmm.DTensor(gpu_ctx, ttype, list(want.shape[::-1]))
    gpu_ctx.op_repeat(up(qmm, gpu_ctx, a), dst)
    gpu_ctx.synchronize()
    assert np.array_equal(dst.numpy().reshape(want.shape), want)


def test_recorded_sequence_replays_and_refuses_a_wait(qmm, gpu_ctx, oracle):
    """b200_graph_begin / _end / _launch: a sequence with a decode mul_mat of the path in it is recorded (nothing runs), replayed on new input and equals the
    same calls issued one by one, bitwise; a call that waits for the device inside the recording makes b200_graph_end report B200_ERR_UNSUPPORTED and
    leaves the context usable."""
    rng = np.random.default_rng(5)
    k, m = 1024, 768
    w = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    qt = qmm.QTensor(gpu_ctx, Q4_0, k, m)
    qt.set(oracle.quantize_weights(Q4_0, w))
    x = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [k])
    h = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [k])
    y = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [m])

    def sequence():
        gpu_ctx.op_norm(x, h, eps=1e-5)
        gpu_ctx.op_scale(h, h, 0.5)
        gpu_ctx.mul_mat_device(qt, h.buf.ptr, 1, y.buf.ptr, 1, 1)
        gpu_ctx.op_unary("gelu", y, y)

    x0 = rng.uniform(-2, 2, k).astype(np.float32)
    x.buf.upload(x0)
    sequence()                                                         # sizes the scratch areas
    gpu_ctx.synchronize()
    y.buf.upload(np.zeros(m, np.float32))
    gpu_ctx.graph_begin()
    sequence()
    g = gpu_ctx.graph_end()
    try:
        assert gpu_ctx.lib.b200_graph_node_count(g) >= 4
        gpu_ctx.synchronize()
        assert not y.numpy().any()                                     # recording executed nothing
        for rep in range(3):
            xi = rng.uniform(-2, 2, k).astype(np.float32)
            x.buf.upload(xi)
            gpu_ctx.graph_launch(g)
            gpu_ctx.synchronize()
            got = y.numpy().copy()
            sequence()
            gpu_ctx.synchronize()
            assert np.array_equal(got, y.numpy())
    finally:
        gpu_ctx.graph_destroy(g)
    gpu_ctx.graph_begin()
    gpu_ctx.op_scale(x, h, 2.0)
    with pytest.raises(qmm.B200Error):
        gpu_ctx.synchronize()                                          # waiting for a stream that is being recorded
    with pytest.raises(qmm.B200Error) as ei:
        gpu_ctx.graph_end()
    assert ei.value.code == qmm.ERR_UNSUPPORTED
    gpu_ctx.op_scale(x, h, 2.0)                                        # the context works again
    gpu_ctx.synchronize()
    assert np.array_equal(h.numpy().reshape(-1), x.numpy().reshape(-1) * 2.0)


@pytest.mark.parametrize("kdt,vdt", [(np.float16, np.float16), (np.float32, np.float32), (np.float16, np.float32)])
@pytest.mark.parametrize("hd,H,N,n_past,n_ctx", [(64, 12, 1, 36, 64), (256, 16, 1, 0, 32), (256, 16, 3, 197, 256), (80, 5, 8, 1016, 1024)])
def test_attention_decode_one_launch(qmm, gpu_ctx, kdt, vdt, hd, H, N, n_past, n_ctx):
    """b200_op_attention_decode against the six operators it stands for (K*Q, SCALE, DIAG_MASK_INF, SOFT_MAX, V*P, merge of the heads) in float64, on
    strided views of a KV cache laid out like GPT-J's (examples/gpt-j/main.cpp:476-512: k rows [n_ctx][n_embd], v transposed [n_embd][n_ctx])"""
    rng = np.random.default_rng(hd + H + N + n_past)
    T, E = n_past + N, hd * H
    kc = rng.uniform(-1, 1, (n_ctx, E)).astype(kdt)                   # cache of k: position-major
    vc = rng.uniform(-1, 1, (E, n_ctx)).astype(vdt)                   # cache of v: one row per embedding dimension
    q = rng.uniform(-1, 1, (N, H, hd)).astype(np.float32)             # [hd][H][N] as computed; read through a permuted view [hd][N][H]
    scale = 1.0 / np.sqrt(hd)
    tk, tv, tq = up(qmm, gpu_ctx, kc), up(qmm, gpu_ctx, vc), up(qmm, gpu_ctx, q)
    ek, ev = kc.itemsize, vc.itemsize
    K = tk.view([hd, T, H, 1], [ek, ek * E, ek * hd, ek * E * n_ctx])
    V = tv.view([T, hd, H, 1], [ev, ev * n_ctx, ev * n_ctx * hd, ev * n_ctx * E])
    Q = tq.view([hd, N, H, 1], [4, 4 * E, 4 * hd, 4 * E * N])
    dst = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [hd, H, N, 1])
    l0 = gpu_ctx.launch_count()
    gpu_ctx.op_attention_decode(Q, K, V, dst, scale, n_past)
    assert gpu_ctx.launch_count() - l0 == 1
    gpu_ctx.synchronize()
    got = dst.numpy().reshape(N, H, hd)
    k64 = kc[:T].astype(np.float64).reshape(T, H, hd)
    v64 = vc[:, :T].astype(np.float64).reshape(H, hd, T)
    want = np.empty((N, H, hd))
    for i in range(N):
        for h in range(H):
            s = (k64[:, h, :] @ q[i, h].astype(np.float64)) * scale
            s[np.arange(T) > n_past + i] = -np.inf
            p = np.exp(s - s.max())
            p /= p.sum()
            want[i, h] = v64[h] @ p
    assert nmse(got, want) <= 1e-10, nmse(got, want)


def test_attention_decode_declines_long_contexts(qmm, gpu_ctx):
    q = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [64, 1, 2, 1])
    k = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [64, 1025, 2, 1])
    v = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [1025, 64, 2, 1])
    d = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [64, 2, 1, 1])
    with pytest.raises(qmm.B200Error) as e:
        gpu_ctx.op_attention_decode(q, k, v, d, 1.0, 1024)
    assert e.value.code == qmm.ERR_UNSUPPORTED


def test_rope_on_strided_rows(qmm, gpu_ctx):
    """ROPE reads and writes through the tensors' strides: q of a fused qkv projection is a column slice (row pitch 3 * n_embd), the result goes to a dense tensor"""
    rng = np.random.default_rng(3)
    hd, H, T = 64, 6, 5
    E = hd * H
    qkv = rng.uniform(-1, 1, (T, 3 * E)).astype(np.float32)
    pos = np.arange(7, 7 + T, dtype=np.int32)
    t = up(qmm, gpu_ctx, qkv)
    k_view = t.view([hd, H, T, 1], [4, 4 * hd, 4 * 3 * E, 4 * 3 * E * T], offset=4 * E)           # the k third of every row
    dst = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [hd, H, T, 1])
    gpu_ctx.op_rope(k_view, up(qmm, gpu_ctx, pos), dst, n_dims=hd, mode=2, n_ctx=64)
    gpu_ctx.synchronize()
    x = qkv[:, E:2 * E].reshape(1, T, H, hd)
    want = rope_numpy(np.ascontiguousarray(x), pos, n_dims=hd, mode=2, n_orig_ctx=0, freq_base=10000.0, freq_scale=1.0, ext_factor=0.0, attn_factor=1.0,
                      beta_fast=0.0, beta_slow=0.0)
    assert nmse(dst.numpy().reshape(1, T, H, hd).astype(np.float64), want.astype(np.float64)) <= 1e-10
