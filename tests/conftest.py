"""Shared fixtures.  `-m "not gpu"` = oracle vs golden vectors / reference, host logic, C-ABI symbols.
`-m gpu` = parity of the CUDA path against the oracle, through the C ABI.  The oracle (oracle/) is the
checker only; the product (ggml-imax_b200/) never calls it."""
import ctypes as C
import importlib.util
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "ggml-imax_b200"
ORACLE_SO = ROOT / "oracle" / "_build" / "libqmm_oracle.so"
REF_SHIM_SO = ROOT / "oracle" / "_ref" / "libref_shim.so"
GOLDEN = ROOT / "tests" / "golden" / "qmm_golden.npz"

Q4_0, Q8_0 = 2, 8
WIRE = {Q4_0: 18, Q8_0: 34}
vp = C.c_void_p


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_qmm():
    """import ggml-imax_b200/qmm.py (the directory name is not an identifier, so go through importlib)."""
    if "ggml_imax_b200_qmm" in sys.modules:
        return sys.modules["ggml_imax_b200_qmm"]
    spec = importlib.util.spec_from_file_location("ggml_imax_b200_qmm", PKG / "qmm.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ggml_imax_b200_qmm"] = mod
    spec.loader.exec_module(mod)
    return mod


class Oracle:
    """ctypes face of oracle/qmm_oracle.c (the CPU restatement of the reference path)."""

    def __init__(self):
        if not ORACLE_SO.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "oracle"])
        self.lib = C.CDLL(str(ORACLE_SO))
        L = self.lib
        L.oracle_fp32_to_fp16.restype = C.c_uint16
        L.oracle_fp32_to_fp16.argtypes = [C.c_float]
        L.oracle_fp16_to_fp32.restype = C.c_float
        L.oracle_fp16_to_fp32.argtypes = [C.c_uint16]
        for name in ("oracle_vec_dot_q4_0_q8_0_scalar", "oracle_vec_dot_q8_0_q8_0_scalar",
                     "oracle_vec_dot_q4_0_q8_0_avx2order", "oracle_vec_dot_q8_0_q8_0_avx2order"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_int64, vp, vp]
        L.oracle_mul_mat.restype = C.c_int
        L.oracle_mul_mat.argtypes = [C.c_int, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_size_t, C.c_size_t, C.c_size_t, vp, C.c_int]
        L.oracle_mul_mat_mt.restype = C.c_int
        L.oracle_mul_mat_mt.argtypes = [C.c_int, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, C.c_int]

    def quantize_row_q8_0(self, x):
        x = np.ascontiguousarray(x, np.float32)
        rows = x.reshape(-1, x.shape[-1])
        k = rows.shape[1]
        out = np.zeros((rows.shape[0], k // 32 * 34), np.uint8)
        for i in range(rows.shape[0]):
            self.lib.oracle_quantize_row_q8_0(rows[i].ctypes.data_as(vp), out[i].ctypes.data_as(vp), C.c_int64(k))
        return out

    def quantize_weights(self, qtype, w):
        """ggml_quantize_chunk(type, ..., imatrix=NULL): the *_reference quantizers, row by row."""
        w = np.ascontiguousarray(w, np.float32)
        rows = w.reshape(-1, w.shape[-1])
        k = rows.shape[1]
        out = np.zeros((rows.shape[0], k // 32 * WIRE[qtype]), np.uint8)
        fn = self.lib.oracle_quantize_row_q4_0_reference if qtype == Q4_0 else self.lib.oracle_quantize_row_q8_0_reference
        for i in range(rows.shape[0]):
            fn(rows[i].ctypes.data_as(vp), out[i].ctypes.data_as(vp), C.c_int64(k))
        return out

    def dequantize(self, qtype, wire, k):
        wire = np.ascontiguousarray(wire, np.uint8).reshape(-1, k // 32 * WIRE[qtype])
        out = np.zeros((wire.shape[0], k), np.float32)
        fn = self.lib.oracle_dequantize_row_q4_0 if qtype == Q4_0 else self.lib.oracle_dequantize_row_q8_0
        for i in range(wire.shape[0]):
            fn(wire[i].ctypes.data_as(vp), out[i].ctypes.data_as(vp), C.c_int64(k))
        return out

    def block_dots(self, qtype, w_wire, q8_wire, k):
        """[n, m, k/32] int32 per-block partial sums."""
        w = np.ascontiguousarray(w_wire, np.uint8).reshape(-1, k // 32 * WIRE[qtype])
        y = np.ascontiguousarray(q8_wire, np.uint8).reshape(-1, k // 32 * 34)
        out = np.zeros((y.shape[0], w.shape[0], k // 32), np.int32)
        fn = self.lib.oracle_block_dots_q4_0_q8_0 if qtype == Q4_0 else self.lib.oracle_block_dots_q8_0_q8_0
        for c in range(y.shape[0]):
            for r in range(w.shape[0]):
                fn(C.c_int64(k), w[r].ctypes.data_as(vp), y[c].ctypes.data_as(vp), out[c, r].ctypes.data_as(vp))
        return out

    def vec_dot(self, qtype, k, x, y, avx2=True):
        name = f"oracle_vec_dot_{'q4_0' if qtype == Q4_0 else 'q8_0'}_q8_0_{'avx2order' if avx2 else 'scalar'}"
        return getattr(self.lib, name)(k, x.ctypes.data_as(vp), y.ctypes.data_as(vp))

    def mul_mat(self, qtype, a_wire, k, m, ne02, ne03, b, avx2=True):
        """a_wire: [ne03, ne02, m, row bytes]; b: [ne13, ne12, n, k] float32 -> dst [ne13, ne12, n, m]."""
        a = np.ascontiguousarray(a_wire, np.uint8)
        b = np.ascontiguousarray(b, np.float32)
        ne13, ne12, n, kk = b.shape
        assert kk == k
        dst = np.zeros((ne13, ne12, n, m), np.float32)
        rc = self.lib.oracle_mul_mat(qtype, a.ctypes.data_as(vp), k, m, ne02, ne03, b.ctypes.data_as(vp), n, ne12, ne13,
                                     k * 4, k * 4 * n, k * 4 * n * ne12, dst.ctypes.data_as(vp), 1 if avx2 else 0)
        assert rc == 0
        return dst

    def mul_mat_mt(self, qtype, a_wire, k, m, b, nthreads):
        a = np.ascontiguousarray(a_wire, np.uint8)
        b = np.ascontiguousarray(b, np.float32)
        n = b.shape[0]
        dst = np.zeros((n, m), np.float32)
        wdata = np.zeros(n * (k // 32) * 34, np.uint8)
        self.lib.oracle_mul_mat_mt(qtype, a.ctypes.data_as(vp), k, m, b.ctypes.data_as(vp), n, dst.ctypes.data_as(vp),
                                   wdata.ctypes.data_as(vp), nthreads)
        return dst


def nmse(a, b):
    """normalized mean squared error exactly as tests/test-backend-ops.cpp:175-188 (a = result, b = reference)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = float(np.sum(b * b))
    num = float(np.sum((a - b) ** 2))
    return num / den if den > 0 else num


MUL_MAT_NMSE_TOL = 5e-4  # tests/test-backend-ops.cpp:921-923
F16_GEMM_NMSE = 1e-6          # our own bound for the fp16 tensor-core prefill path (operands rounded once to fp16; measured ~1e-7)


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def qmm():
    return load_qmm()


@pytest.fixture(scope="session")
def gpu_ctx(qmm):
    """A context on cuda:0.  Fails (does not skip) when the extension or the GPU is missing: -m gpu runs
    must exercise the native path."""
    ctx = qmm.Context(0)
    yield ctx
    ctx.close()
