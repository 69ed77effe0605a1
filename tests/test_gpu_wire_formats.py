"""Q5_0 and IQ4_NL -- the sibling 32-element block formats that share the Q8_0 activation path -- through the C ABI against
the UNMODIFIED reference CPU backend (oracle/_ref/libref_shim.so: ggml_quantize_chunk for the weights, ggml_backend_graph_compute of one
MUL_MAT node, dequantize_row for GET_ROWS) on the same inputs.  These tensors stay in wire format on the device.  MUL_MAT within
test-backend-ops' NMSE <= 5e-4 (tests/test-backend-ops.cpp:921-923; measured ~1e-13: same integer dots, same quantize_row_q8_0, only the
summation order differs), GET_ROWS bit-identical to the reference's dequantize_row_*."""
import ctypes as C

import numpy as np
import pytest

from conftest import REF_SHIM_SO, MUL_MAT_NMSE_TOL, nmse

pytestmark = pytest.mark.gpu
vp = C.c_void_p
Q5_0, IQ4_NL = 6, 20


@pytest.fixture(scope="module")
def ref():
    assert REF_SHIM_SO.exists(), f"{REF_SHIM_SO} must be prebuilt (make -C oracle ref) and travel with the snapshot"
    lib = C.CDLL(str(REF_SHIM_SO))
    lib.ref_quantize_chunk.restype = C.c_size_t
    lib.ref_quantize_chunk.argtypes = [C.c_int, vp, vp, C.c_int64, C.c_int64]
    lib.ref_row_size.restype = C.c_size_t
    lib.ref_row_size.argtypes = [C.c_int, C.c_int64]
    lib.ref_to_float.argtypes = [C.c_int, vp, vp, C.c_int64]
    lib.ref_mm_create.restype = vp
    lib.ref_mm_create.argtypes = [C.c_int] + [C.c_int64] * 7 + [C.c_int]
    for f in ("ref_mm_set_a", "ref_mm_set_b", "ref_mm_get_out"):
        getattr(lib, f).argtypes = [vp, vp]
    lib.ref_mm_compute.restype = C.c_double
    lib.ref_mm_compute.argtypes = [vp, C.c_int]
    lib.ref_mm_free.argtypes = [vp]
    lib.ref_time_init()
    return lib


def quantize(ref, qtype, w):
    m, k = w.shape
    out = np.zeros(m * ref.ref_row_size(qtype, k), np.uint8)
    ref.ref_quantize_chunk(qtype, w.ctypes.data_as(vp), out.ctypes.data_as(vp), m, k)
    return out


@pytest.mark.parametrize("qtype", [Q5_0, IQ4_NL])
@pytest.mark.parametrize("m,k,n,ne02,nr2", [(300, 4096, 1, 1, 1), (64, 1024, 5, 1, 1), (17, 256, 33, 1, 1), (16, 256, 3, 4, 2), (1000, 768, 8, 1, 1)])
def test_mul_mat_against_the_reference_cpu_backend(qmm, gpu_ctx, ref, qtype, m, k, n, ne02, nr2):
    rng = np.random.default_rng(qtype + m + k + n)
    w = rng.uniform(-1, 1, (ne02 * m, k)).astype(np.float32)
    wire = quantize(ref, qtype, w)
    x = rng.uniform(-1, 1, (ne02 * nr2, n, k)).astype(np.float32)
    h = vp(ref.ref_mm_create(qtype, k, m, ne02, 1, n, ne02 * nr2, 1, 4))
    ref.ref_mm_set_a(h, wire.ctypes.data_as(vp))
    ref.ref_mm_set_b(h, x.ctypes.data_as(vp))
    ref.ref_mm_compute(h, 1)
    want = np.zeros((ne02 * nr2, n, m), np.float32)
    ref.ref_mm_get_out(h, want.ctypes.data_as(vp))
    ref.ref_mm_free(h)
    wd = gpu_ctx.alloc(wire.nbytes)
    wd.upload(wire)
    xd = gpu_ctx.to_device(x)
    yd = gpu_ctx.alloc(want.nbytes)
    a = qmm.MulMatArgs()
    a.type = qtype
    a.src0_dev = wd.ptr
    a.src0_nblocks_total = ne02 * m * (k // 32)
    a.ne00, a.ne01, a.ne02, a.ne03 = k, m, ne02, 1
    a.src1_dev = xd.ptr
    a.ne11, a.ne12, a.ne13 = n, ne02 * nr2, 1
    a.nb11, a.nb12, a.nb13 = k * 4, k * 4 * n, k * 4 * n * ne02 * nr2
    a.dst_dev = yd.ptr
    gpu_ctx._check(gpu_ctx.lib.b200_mul_mat(gpu_ctx.h, C.byref(a)))
    gpu_ctx.synchronize()
    got = yd.download(np.float32, want.size).reshape(want.shape)
    assert np.isfinite(got).all()
    e = nmse(got, want)
    assert e <= MUL_MAT_NMSE_TOL and e <= 1e-10, e


@pytest.mark.parametrize("qtype", [Q5_0, IQ4_NL])
def test_get_rows_is_the_reference_dequantization(qmm, gpu_ctx, ref, qtype):
    rng = np.random.default_rng(qtype)
    m, k, r = 200, 768, 31
    w = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    wire = quantize(ref, qtype, w)
    deq = np.zeros((m, k), np.float32)
    ref.ref_to_float(qtype, wire.ctypes.data_as(vp), deq.ctypes.data_as(vp), m * k)
    rows = rng.integers(0, m, (1, r)).astype(np.int32)
    wd = gpu_ctx.alloc(wire.nbytes)
    wd.upload(wire)
    rsz = ref.ref_row_size(qtype, k)
    src = qmm.DTensor.__new__(qmm.DTensor)
    t = qmm.Tensor()
    t.data, t.type = wd.ptr, qtype
    t.ne = (C.c_int64 * 4)(k, m, 1, 1)
    t.nb = (C.c_int64 * 4)(rsz // (k // 32), rsz, rsz * m, rsz * m)

    class Raw:
        def desc(self):
            return t
    dst = qmm.DTensor(gpu_ctx, qmm.TYPE_F32, [k, r, 1, 1])
    gpu_ctx.op_get_rows(Raw(), qmm.DTensor.from_numpy(gpu_ctx, rows), dst)
    gpu_ctx.synchronize()
    assert np.array_equal(dst.numpy().reshape(r, k), deq[rows[0]])
