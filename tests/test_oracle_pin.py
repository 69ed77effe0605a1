"""Pin oracle/qmm_oracle.c: (a) against golden vectors produced by the unmodified reference
(tests/golden/qmm_golden.npz, made by oracle/make_golden.py) and (b) against the reference itself
(oracle/_ref/libref_shim.so) whenever that prebuilt file is present.  All checks are bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import Q4_0, Q8_0, WIRE, REF_SHIM_SO, nmse, vp

NAMES = {Q4_0: "q4_0", Q8_0: "q8_0"}


def test_fp16_roundtrip_exhaustive(oracle):
    hs = np.arange(65536, dtype=np.uint16)
    f = hs.view(np.float16).astype(np.float32)
    for h in range(0, 65536, 3):
        v = np.float32(oracle.lib.oracle_fp16_to_fp32(h))
        assert v.tobytes() == f[h].tobytes() or (np.isnan(v) and np.isnan(f[h]))
        if np.isfinite(f[h]):
            assert oracle.lib.oracle_fp32_to_fp16(float(f[h])) == h


def test_fp32_to_fp16_rne(oracle):
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-70000, 70000, 5000), rng.uniform(-1e-4, 1e-4, 5000), rng.uniform(-1e-7, 1e-7, 5000),
                         [65504.0, 65519.9, 65520.0, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0000001, 0.0, -0.0]]).astype(np.float32)
    with np.errstate(over="ignore"):
        ref = xs.astype(np.float16).view(np.uint16)
    got = np.array([oracle.lib.oracle_fp32_to_fp16(float(v)) for v in xs], dtype=np.uint16)
    assert np.array_equal(got, ref)


def test_quantize_row_q8_0_golden(oracle, golden):
    got = oracle.quantize_row_q8_0(golden["act_x"])
    assert np.array_equal(got, golden["act_q8_0"])


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_weight_quantizers_golden(oracle, golden, qtype):
    got = oracle.quantize_weights(qtype, golden["w_f32"])
    assert np.array_equal(got, golden[f"w_{NAMES[qtype]}"])
    deq = oracle.dequantize(qtype, got, golden["w_f32"].shape[1])
    assert np.array_equal(deq.view(np.uint32), golden[f"w_{NAMES[qtype]}_dequant"].view(np.uint32))


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_vec_dot_golden_bit_exact(oracle, golden, qtype):
    w, y = golden[f"w_{NAMES[qtype]}"], golden["act_q8_0"]
    k = golden["act_x"].shape[1]
    ref = golden[f"vec_dot_{NAMES[qtype]}"]
    for i in range(w.shape[0]):
        for j in range(y.shape[0]):
            got = np.float32(oracle.vec_dot(qtype, k, w[i], y[j], avx2=True))
            assert got.tobytes() == ref[i, j].tobytes(), (i, j, got, ref[i, j])


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_block_dots_sum_to_vec_dot(oracle, golden, qtype):
    """sum_b float(dot_b) * d_w * d_x (scalar order) must equal the scalar vec_dot: ties the int32 block
    partials to the float result."""
    w, y = golden[f"w_{NAMES[qtype]}"], golden["act_q8_0"]
    k = golden["act_x"].shape[1]
    dots = oracle.block_dots(qtype, w, y, k)
    wb = WIRE[qtype]
    np.seterr(invalid="ignore", over="ignore")   # the 1e30 row overflows fp16 d -> inf/nan, compared bit-wise anyway
    for r in range(w.shape[0]):
        dw = w[r].reshape(-1, wb)[:, :2].copy().view(np.float16).astype(np.float32).ravel()
        for c in range(y.shape[0]):
            dy = y[c].reshape(-1, 34)[:, :2].copy().view(np.float16).astype(np.float32).ravel()
            s = np.float32(0)
            for b in range(k // 32):
                if qtype == Q4_0:
                    s = np.float32(s + np.float32(np.float32(np.float32(dots[c, r, b]) * dw[b]) * dy[b]))
                else:
                    s = np.float32(s + np.float32(np.float32(dots[c, r, b]) * np.float32(dw[b] * dy[b])))
            got = np.float32(oracle.vec_dot(qtype, k, w[r], y[c], avx2=False))
            assert s.tobytes() == got.tobytes()


def test_mul_mat_golden_bit_exact(oracle, golden):
    cases = golden["mm_cases"]
    for ci, (t, m, n, k, bs0, bs1, nr0, nr1) in enumerate(cases.tolist()):
        a = golden[f"mm{ci}_a"]
        b = golden[f"mm{ci}_b"].astype(np.float32)
        ref = golden[f"mm{ci}_out"]
        got = oracle.mul_mat(t, a, k, m, bs0, bs1, b, avx2=True)
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"case {ci}: nmse {nmse(got, ref)}"
        # the scalar summation order differs only by fp32 rounding
        got_s = oracle.mul_mat(t, a, k, m, bs0, bs1, b, avx2=False)
        assert nmse(got_s, ref) < 1e-7


def test_mul_mat_rejects_bad_shapes(oracle):
    a = np.zeros(18, np.uint8)
    b = np.zeros((1, 1, 1, 32), np.float32)
    d = np.zeros(4, np.float32)
    L = oracle.lib
    assert L.oracle_mul_mat(Q4_0, a.ctypes.data_as(vp), 33, 1, 1, 1, b.ctypes.data_as(vp), 1, 1, 1, 128, 128, 128, d.ctypes.data_as(vp), 1) == -1
    assert L.oracle_mul_mat(Q4_0, a.ctypes.data_as(vp), 32, 1, 2, 1, b.ctypes.data_as(vp), 1, 3, 1, 128, 128, 128, d.ctypes.data_as(vp), 1) == -1
    assert L.oracle_mul_mat(1, a.ctypes.data_as(vp), 32, 1, 1, 1, b.ctypes.data_as(vp), 1, 1, 1, 128, 128, 128, d.ctypes.data_as(vp), 1) == -1


def test_mt_variant_matches(oracle):
    rng = np.random.default_rng(5)
    k, m, n = 128, 37, 3
    w = oracle.quantize_weights(Q4_0, rng.uniform(-1, 1, (m, k)))
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    ref = oracle.mul_mat(Q4_0, w, k, m, 1, 1, x[None, None], avx2=True)[0, 0]
    for t in (1, 3, 8):
        assert np.array_equal(oracle.mul_mat_mt(Q4_0, w, k, m, x, t), ref)


@pytest.mark.skipif(not REF_SHIM_SO.exists(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_against_live_reference(oracle):
    """Fresh random inputs through the unmodified reference vs the restatement, bit for bit."""
    r = C.CDLL(str(REF_SHIM_SO))
    r.ref_vec_dot.restype = C.c_float
    r.ref_time_init()
    rng = np.random.default_rng(77)
    k, nrows = 1024, 300
    x = (rng.standard_normal((nrows, k)) * rng.uniform(1e-3, 1e3, (nrows, 1))).astype(np.float32)
    ref = np.zeros((nrows, k // 32 * 34), np.uint8)
    for i in range(nrows):
        r.ref_from_float(Q8_0, x[i].ctypes.data_as(vp), ref[i].ctypes.data_as(vp), C.c_int64(k))
    assert np.array_equal(oracle.quantize_row_q8_0(x), ref)
    for qtype in (Q4_0, Q8_0):
        w = oracle.quantize_weights(qtype, rng.uniform(-1, 1, (64, k)))
        for i in range(64):
            a = np.float32(oracle.vec_dot(qtype, k, w[i], ref[i]))
            b = np.float32(r.ref_vec_dot(qtype, C.c_int64(k), w[i].ctypes.data_as(vp), ref[i].ctypes.data_as(vp)))
            assert a.tobytes() == b.tobytes()
