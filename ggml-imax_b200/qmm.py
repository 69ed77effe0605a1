"""ctypes binding of include/ggml_b200.h -- the host-side mirror used by tests/, bench.py and smoke().

Names and argument meaning follow the reference path they replace:
  quantize_row_q8_0      src/ggml-quants.c:465           -> Context.quantize_row_q8_0
  ggml_backend_tensor_set/get on a Q4_0/Q8_0 tensor (src/ggml-backend.c:221-247) -> QTensor.set / QTensor.get
  ggml_compute_forward_mul_mat (src/ggml.c:11808)        -> Context.mul_mat

Python is plumbing only: every byte of arithmetic happens in the CUDA kernels behind the C ABI.  If the
shared library is missing or no sm_100 device is usable this module raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libggml_b200.so"
BACKEND_LIB_PATH = PKG_DIR / "lib" / "libggml-b200-backend.so"

TYPE_F32, TYPE_F16, TYPE_Q4_0, TYPE_Q8_0, TYPE_I16, TYPE_I32 = 0, 1, 2, 8, 25, 26
TYPE_Q5_0, TYPE_IQ4_NL = 6, 20      # kept in wire format on the device (b200_wire_formats.cu)
OP_ADD, OP_MUL, OP_DIV = 0, 1, 2
UNARY = {n: i for i, n in enumerate(["abs", "sgn", "neg", "step", "tanh", "elu", "relu", "sigmoid", "gelu", "gelu_quick", "silu", "hardswish",
                                     "hardsigmoid"])}          # enum ggml_unary_op
QK = 32
WIRE_BYTES = {TYPE_Q4_0: 18, TYPE_Q8_0: 34}
TYPE_NAMES = {TYPE_Q4_0: "q4_0", TYPE_Q8_0: "q8_0"}
MM_FORCE_GEMV, MM_FORCE_GEMM, MM_EXPORT = 1, 2, 4

OK, ERR_CUDA, ERR_INVALID, ERR_UNSUPPORTED, ERR_ALLOC = 0, -1, -2, -3, -4


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ggml_b200 error {code}: {msg}")
        self.code = code


class MulMatArgs(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("flags", C.c_int32),
        ("src0_dev", C.c_void_p), ("src0_nblocks_total", C.c_int64), ("src0_block_off", C.c_int64),
        ("ne00", C.c_int64), ("ne01", C.c_int64), ("ne02", C.c_int64), ("ne03", C.c_int64),
        ("src1_dev", C.c_void_p),
        ("ne11", C.c_int64), ("ne12", C.c_int64), ("ne13", C.c_int64),
        ("nb11", C.c_size_t), ("nb12", C.c_size_t), ("nb13", C.c_size_t),
        ("dst_dev", C.c_void_p),
    ]


class Tensor(C.Structure):
    """b200_tensor: what a glue-operator kernel needs of a ggml_tensor (device address, type, ne[], nb[] in bytes)."""
    _fields_ = [("data", C.c_void_p), ("type", C.c_int32), ("reserved", C.c_int32), ("ne", C.c_int64 * 4), ("nb", C.c_int64 * 4),
                ("q_total_blocks", C.c_int64), ("q_block_off", C.c_int64)]


class RopeParams(C.Structure):
    """b200_rope_params: the op_params of GGML_OP_ROPE (src/ggml.c:5866-5889)"""
    _fields_ = [("n_dims", C.c_int32), ("mode", C.c_int32), ("n_ctx", C.c_int32), ("n_orig_ctx", C.c_int32), ("freq_base", C.c_float),
                ("freq_scale", C.c_float), ("ext_factor", C.c_float), ("attn_factor", C.c_float), ("beta_fast", C.c_float), ("beta_slow", C.c_float),
                ("xpos_base", C.c_float), ("xpos_down", C.c_int32)]


class Epilogue(C.Structure):
    _fields_ = [("bias_dev", C.c_void_p), ("residual_dev", C.c_void_p), ("act", C.c_int32), ("reserved", C.c_int32), ("residual2_dev", C.c_void_p)]


EPI_NONE, EPI_GELU = 0, 1
MAX_RANKS = 8


class Gather(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("slot", C.c_int32), ("wait_slot", C.c_int32), ("row0", C.c_int64),
        ("peer_dst", C.c_void_p * MAX_RANKS), ("state", C.c_void_p),
    ]


class PlanSplit(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("peer_arena", C.c_void_p * MAX_RANKS),
        ("row0", C.POINTER(C.c_int64)), ("m_total", C.POINTER(C.c_int64)),
    ]


# every symbol include/ggml_b200.h declares (tests check the list against the header and the .so)
_SIGNATURES = {
    "b200_device_count": (C.c_int, []),
    "b200_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "b200_ctx_create_on_stream": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200_ctx_destroy": (None, [C.c_void_p]),
    "b200_last_error": (C.c_char_p, [C.c_void_p]),
    "b200_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "b200_ctx_device": (C.c_int, [C.c_void_p]),
    "b200_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "b200_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "b200_ctx_set_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "b200_malloc": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t]),
    "b200_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_memset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "b200_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200_upload_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200_download_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200_copy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200_copy_2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "b200_enable_peer_access": (C.c_int, [C.c_void_p, C.c_int]),
    "b200_synchronize": (C.c_int, [C.c_void_p]),
    "b200_event_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200_event_destroy": (None, [C.c_void_p]),
    "b200_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_event_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_event_synchronize": (C.c_int, [C.c_void_p]),
    "b200_host_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "b200_host_free": (C.c_int, [C.c_void_p]),
    "b200_graph_begin": (C.c_int, [C.c_void_p]),
    "b200_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_graph_destroy": (None, [C.c_void_p]),
    "b200_reserve_workspace": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64]),
    "b200_set_quantized": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]),
    "b200_get_quantized": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]),
    "b200_repack_from_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]),
    "b200_unrepack_to_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]),
    "b200_quantize_q8_0": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_size_t, C.c_void_p, C.c_void_p]),
    "b200_quantize_q8_0_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_size_t, C.c_void_p]),
    "b200_mul_mat": (C.c_int, [C.c_void_p, C.POINTER(MulMatArgs)]),
    "b200_mul_mat_fused": (C.c_int, [C.c_void_p, C.POINTER(MulMatArgs), C.POINTER(Epilogue)]),
    "b200_mul_mat_batch": (C.c_int, [C.c_void_p, C.POINTER(MulMatArgs), C.c_int]),
    "b200_mul_mat_gather": (C.c_int, [C.c_void_p, C.POINTER(MulMatArgs), C.POINTER(Gather)]),
    "b200_mul_mat_gather_batch": (C.c_int, [C.c_void_p, C.POINTER(MulMatArgs), C.POINTER(Gather), C.c_int]),
    "b200_gather_finish": (C.c_int, [C.c_void_p, C.POINTER(Gather), C.c_void_p, C.c_void_p, C.c_int64]),
    "b200_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200_ipc_import": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_plan_arena_bytes": (C.c_size_t, [C.POINTER(MulMatArgs), C.c_int, C.POINTER(PlanSplit)]),
    "b200_plan_create": (C.c_int, [C.c_void_p, C.POINTER(MulMatArgs), C.c_int, C.POINTER(PlanSplit), C.POINTER(C.c_void_p)]),
    "b200_plan_analyze": (C.c_int, [C.POINTER(MulMatArgs), C.c_int, C.POINTER(PlanSplit), C.POINTER(C.c_int32)]),
    "b200_plan_published": (C.c_int, [C.POINTER(MulMatArgs), C.c_int, C.POINTER(PlanSplit), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "b200_plan_plain_stores": (C.c_int, [C.POINTER(MulMatArgs), C.c_int, C.POINTER(PlanSplit), C.POINTER(C.c_int32)]),
    "b200_plan_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_plan_destroy": (None, [C.c_void_p]),
    "b200_plan_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200_op_get_rows": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor)]),
    "b200_op_binary": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor)]),
    "b200_op_unary": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor)]),
    "b200_op_norm": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.c_float, C.c_int]),
    "b200_op_scale": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.c_float]),
    "b200_op_diag_mask_inf": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.c_int]),
    "b200_op_soft_max": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.c_float, C.c_float, C.c_int]),
    "b200_op_copy": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor)]),
    "b200_graph_node_count": (C.c_int64, [C.c_void_p]),
    "b200_op_rope": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(RopeParams)]),
    "b200_op_repeat": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor)]),
    "b200_op_attention_decode": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.c_float, C.c_int]),
    "b200_op_mul_mat_dense": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor)]),
    "b200_block_dots": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "b200_mul_mat_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
}

_lib = None


def declared_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def load_library() -> C.CDLL:
    """dlopen the in-tree C-ABI library; loud failure if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} not built: run `make -C {PKG_DIR}` (or __graft_entry__.build()); "
                           "the B200 path has no CPU fallback")
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def device_count() -> int:
    return load_library().b200_device_count()


def _ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, DeviceBuffer):
        return C.c_void_p(a.ptr)
    return C.c_void_p(int(a))


class DeviceBuffer:
    """Device allocation owned by a Context (b200_malloc / b200_free)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        ctx._check(ctx.lib.b200_malloc(ctx.h, C.byref(p), self.nbytes))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.ctx._check(self.ctx.lib.b200_free(self.ctx.h, C.c_void_p(self.ptr)))
            self.ptr = 0

    def upload(self, arr: np.ndarray, offset: int = 0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.b200_upload(self.ctx.h, C.c_void_p(self.ptr + offset), _ptr(arr), arr.nbytes))

    def download(self, dtype, count: int, offset: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.b200_download(self.ctx.h, _ptr(out), C.c_void_p(self.ptr + offset), out.nbytes))
        return out


_NP_TYPES = {TYPE_F32: np.float32, TYPE_F16: np.float16, TYPE_I16: np.int16, TYPE_I32: np.int32}


class DTensor:
    """A dense tensor on the device for the glue operators: ggml shape order (ne[0] contiguous), optional byte strides / offset
    into a parent buffer (views, permutes).  from_numpy takes an array whose LAST axis is ne[0]."""

    def __init__(self, ctx: "Context", ttype: int, ne, nb=None, buf: DeviceBuffer | None = None, offset: int = 0):
        ne = list(ne) + [1] * (4 - len(ne))
        es = np.dtype(_NP_TYPES[ttype]).itemsize
        if nb is None:
            nb = [es, es * ne[0], es * ne[0] * ne[1], es * ne[0] * ne[1] * ne[2]]
        self.ctx, self.type, self.ne, self.nb, self.offset = ctx, ttype, ne, list(nb), offset
        self.buf = buf if buf is not None else ctx.alloc(max(16, es * ne[0] * ne[1] * ne[2] * ne[3]))

    @classmethod
    def from_numpy(cls, ctx: "Context", arr: np.ndarray) -> "DTensor":
        ttype = {np.dtype(v): k for k, v in _NP_TYPES.items()}[arr.dtype]
        t = cls(ctx, ttype, list(arr.shape[::-1]))
        t.buf.upload(arr)
        return t

    def view(self, ne, nb, offset: int = 0) -> "DTensor":
        return DTensor(self.ctx, self.type, ne, nb, self.buf, self.offset + offset)

    def desc(self) -> Tensor:
        t = Tensor()
        t.data = self.buf.ptr + self.offset
        t.type = self.type
        t.ne = (C.c_int64 * 4)(*self.ne)
        t.nb = (C.c_int64 * 4)(*self.nb)
        return t

    def numpy(self) -> np.ndarray:
        """the dense contents (only for tensors that own their whole buffer contiguously)"""
        n = self.ne[0] * self.ne[1] * self.ne[2] * self.ne[3]
        return self.buf.download(_NP_TYPES[self.type], n, self.offset).reshape(self.ne[::-1])


class QTensor:
    """A Q4_0/Q8_0 tensor [k, m, ne02, ne03] resident on the device in the repacked plane layout.

    set()/get() are ggml_backend_tensor_set/get for a quantized tensor: wire-format blocks in, wire-format out."""

    def __init__(self, ctx: "Context", qtype: int, k: int, m: int, ne02: int = 1, ne03: int = 1, buf: DeviceBuffer | None = None,
                 ptr: int | None = None):
        assert qtype in WIRE_BYTES and k % QK == 0
        self.ctx, self.type, self.k, self.m, self.ne02, self.ne03 = ctx, qtype, k, m, ne02, ne03
        self.nblocks = (k // QK) * m * ne02 * ne03
        self.nbytes = self.nblocks * WIRE_BYTES[qtype]
        self.buf = None
        if ptr is not None:
            self.ptr = int(ptr)
        else:
            self.buf = buf or DeviceBuffer(ctx, self.nbytes)
            self.ptr = self.buf.ptr

    def set(self, wire: np.ndarray, block_off: int = 0):
        wire = np.ascontiguousarray(wire).view(np.uint8).reshape(-1)
        nb = wire.nbytes // WIRE_BYTES[self.type]
        assert nb * WIRE_BYTES[self.type] == wire.nbytes
        self.ctx._check(self.ctx.lib.b200_set_quantized(self.ctx.h, self.type, C.c_void_p(self.ptr), self.nblocks, _ptr(wire),
                                                        block_off, nb))

    def get(self, block_off: int = 0, nblocks: int | None = None) -> np.ndarray:
        nb = self.nblocks - block_off if nblocks is None else nblocks
        out = np.empty(nb * WIRE_BYTES[self.type], dtype=np.uint8)
        self.ctx._check(self.ctx.lib.b200_get_quantized(self.ctx.h, self.type, C.c_void_p(self.ptr), self.nblocks, _ptr(out),
                                                        block_off, nb))
        return out

    def desc(self) -> Tensor:
        """as a b200_tensor (wire strides, repacked root addressing) for b200_op_get_rows"""
        w = WIRE_BYTES[self.type]
        t = Tensor()
        t.data = self.ptr
        t.type = self.type
        t.ne = (C.c_int64 * 4)(self.k, self.m, self.ne02, self.ne03)
        nb1 = w * (self.k // QK)
        t.nb = (C.c_int64 * 4)(w, nb1, nb1 * self.m, nb1 * self.m * self.ne02)
        t.q_total_blocks = self.nblocks
        t.q_block_off = 0
        return t

    def free(self):
        if self.buf is not None:
            self.buf.free()


class Context:
    """One (device, stream): b200_ctx.  `stream` = a raw cudaStream_t (int) to borrow, e.g. torch's current stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        if stream is None:
            rc = self.lib.b200_ctx_create(device, C.byref(h))
        else:
            rc = self.lib.b200_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h))
        if rc != OK:
            raise B200Error(rc, (self.lib.b200_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.b200_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != OK:
            raise B200Error(rc, (self.lib.b200_last_error(self.h) or b"").decode())

    # -- the operators either side of the path (b200_ops.cu); every argument a DTensor / QTensor, asynchronous
    @staticmethod
    def _d(t):
        return None if t is None else C.byref(t.desc())

    def op_get_rows(self, src0, rows, dst):
        self._check(self.lib.b200_op_get_rows(self.h, self._d(src0), self._d(rows), self._d(dst)))

    def op_binary(self, op: int, a, b, dst):
        self._check(self.lib.b200_op_binary(self.h, op, self._d(a), self._d(b), self._d(dst)))

    def op_unary(self, name: str, a, dst):
        self._check(self.lib.b200_op_unary(self.h, UNARY[name], self._d(a), self._d(dst)))

    def op_norm(self, a, dst, eps: float, gain=None, bias=None, rms: bool = False):
        self._check(self.lib.b200_op_norm(self.h, self._d(a), self._d(gain), self._d(bias), self._d(dst), C.c_float(eps), int(rms)))

    def op_scale(self, a, dst, s: float):
        self._check(self.lib.b200_op_scale(self.h, self._d(a), self._d(dst), C.c_float(s)))

    def op_diag_mask_inf(self, a, dst, n_past: int):
        self._check(self.lib.b200_op_diag_mask_inf(self.h, self._d(a), self._d(dst), n_past))

    def op_soft_max(self, a, dst, mask=None, scale: float = 1.0, max_bias: float = 0.0, n_past: int = -1):
        self._check(self.lib.b200_op_soft_max(self.h, self._d(a), self._d(mask), self._d(dst), C.c_float(scale), C.c_float(max_bias), n_past))

    def op_copy(self, a, dst):
        self._check(self.lib.b200_op_copy(self.h, self._d(a), self._d(dst)))

    def op_rope(self, a, pos, dst, n_dims: int, mode: int = 0, n_ctx: int = 0, n_orig_ctx: int = 0, freq_base: float = 10000.0, freq_scale: float = 1.0,
                ext_factor: float = 0.0, attn_factor: float = 1.0, beta_fast: float = 0.0, beta_slow: float = 0.0, xpos_base: float = 0.0, xpos_down: bool = False):
        rp = RopeParams(n_dims, mode, n_ctx, n_orig_ctx, freq_base, freq_scale, ext_factor, attn_factor, beta_fast, beta_slow, xpos_base, int(xpos_down))
        self._check(self.lib.b200_op_rope(self.h, self._d(a), self._d(pos), self._d(dst), C.byref(rp)))

    def op_attention_decode(self, q, k, v, dst, scale: float, n_past: int = -1):
        self._check(self.lib.b200_op_attention_decode(self.h, self._d(q), self._d(k), self._d(v), self._d(dst), C.c_float(scale), n_past))

    def op_repeat(self, a, dst):
        self._check(self.lib.b200_op_repeat(self.h, self._d(a), self._d(dst)))

    def op_mul_mat_dense(self, a, b, dst):
        self._check(self.lib.b200_op_mul_mat_dense(self.h, self._d(a), self._d(b), self._d(dst)))

    # -- plumbing
    def set_option(self, key: str, value: int):
        self._check(self.lib.b200_ctx_set_option(self.h, key.encode(), int(value)))

    def launch_count(self) -> int:
        return int(self.lib.b200_ctx_launch_count(self.h))

    def stream(self) -> int:
        return int(self.lib.b200_ctx_stream(self.h) or 0)

    def synchronize(self):
        self._check(self.lib.b200_synchronize(self.h))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr: np.ndarray) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        b = DeviceBuffer(self, max(arr.nbytes, 1))
        if arr.nbytes:
            b.upload(arr)
        return b

    def set_trace(self, buf, max_launches: int):
        self._check(self.lib.b200_ctx_set_trace(self.h, _ptr(buf), max_launches))

    def graph_begin(self):
        self._check(self.lib.b200_graph_begin(self.h))

    def graph_end(self) -> int:
        g = C.c_void_p()
        self._check(self.lib.b200_graph_end(self.h, C.byref(g)))
        return g.value

    def graph_launch(self, g: int):
        self._check(self.lib.b200_graph_launch(self.h, C.c_void_p(g)))

    def graph_destroy(self, g: int):
        self.lib.b200_graph_destroy(C.c_void_p(g))

    def reserve_workspace(self, qtype: int, k: int, m: int, n: int):
        self._check(self.lib.b200_reserve_workspace(self.h, qtype, k, m, n))

    # -- the path
    def quantize_row_q8_0(self, x: np.ndarray) -> np.ndarray:
        """x: [nrows, k] float32 (host).  Returns nrows*(k/32) block_q8_0 records (uint8 [nrows, k/32*34]) made on the GPU."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x[None, :]
        nrows, k = x.shape
        xd = self.to_device(x)
        out = self.alloc(max(nrows * (k // QK) * 34, 1))
        try:
            self._check(self.lib.b200_quantize_q8_0_blocks(self.h, C.c_void_p(xd.ptr), k, nrows, k * 4, C.c_void_p(out.ptr)))
            return out.download(np.uint8, nrows * (k // QK) * 34).reshape(nrows, -1)
        finally:
            xd.free()
            out.free()

    def quantize_q8_0_planar(self, x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        x = np.ascontiguousarray(x, dtype=np.float32)
        nrows, k = x.shape
        xd = self.to_device(x)
        qs = self.alloc(nrows * k)
        d = self.alloc(nrows * (k // QK) * 2)
        try:
            self._check(self.lib.b200_quantize_q8_0(self.h, C.c_void_p(xd.ptr), k, nrows, k * 4, C.c_void_p(qs.ptr), C.c_void_p(d.ptr)))
            return qs.download(np.int8, nrows * k).reshape(nrows, k), d.download(np.uint16, nrows * (k // QK)).reshape(nrows, -1)
        finally:
            xd.free(); qs.free(); d.free()

    def mul_mat_device(self, w: QTensor, x_ptr: int, n: int, dst_ptr: int, ne12: int = 1, ne13: int = 1,
                       nb11: int | None = None, nb12: int | None = None, nb13: int | None = None, flags: int = 0,
                       block_off: int = 0, m: int | None = None, nblocks_total: int | None = None):
        """Asynchronous mul_mat on device pointers (what graph_compute does for one MUL_MAT node)."""
        a = MulMatArgs()
        a.type, a.flags = w.type, flags
        a.src0_dev = w.ptr
        a.src0_nblocks_total = w.nblocks if nblocks_total is None else nblocks_total
        a.src0_block_off = block_off
        a.ne00, a.ne01, a.ne02, a.ne03 = w.k, (w.m if m is None else m), w.ne02, w.ne03
        a.src1_dev = x_ptr
        a.ne11, a.ne12, a.ne13 = n, ne12, ne13
        a.nb11 = w.k * 4 if nb11 is None else nb11
        a.nb12 = a.nb11 * n if nb12 is None else nb12
        a.nb13 = a.nb12 * ne12 if nb13 is None else nb13
        a.dst_dev = dst_ptr
        self._check(self.lib.b200_mul_mat(self.h, C.byref(a)))

    def ipc_export(self, ptr: int) -> bytes:
        h = (C.c_ubyte * 64)()
        self._check(self.lib.b200_ipc_export(self.h, C.c_void_p(ptr), h))
        return bytes(h)

    def ipc_import(self, handle: bytes) -> int:
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self.lib.b200_ipc_import(self.h, h, C.byref(p)))
        return p.value

    def mul_mat_gather(self, w: QTensor, x_ptr: int, gather: "Gather", m: int | None = None):
        """decode mul_mat of this rank's row slice with the all-gather fused into the epilogue (b200_mul_mat_gather)"""
        a = MulMatArgs()
        a.type = w.type
        a.src0_dev = w.ptr
        a.src0_nblocks_total = w.nblocks
        a.ne00, a.ne01, a.ne02, a.ne03 = w.k, (w.m if m is None else m), 1, 1
        a.src1_dev = x_ptr
        a.ne11, a.ne12, a.ne13 = 1, 1, 1
        a.nb11 = a.nb12 = a.nb13 = w.k * 4
        a.dst_dev = gather.peer_dst[gather.rank]   # unused: results leave as LL elements
        self._check(self.lib.b200_mul_mat_gather(self.h, C.byref(a), C.byref(gather)))

    def mul_mat_gather_batch(self, items):
        """items: [(QTensor slice, x_ptr, Gather, rows)] -- independent same-input slices, one launch when possible"""
        args = (MulMatArgs * len(items))()
        gs = (Gather * len(items))()
        for j, (w, x_ptr, g, m) in enumerate(items):
            a = self.make_args(w, x_ptr, 1, g.peer_dst[g.rank], m=m)
            a.ne02 = a.ne03 = 1
            args[j] = a
            gs[j] = g
        self._check(self.lib.b200_mul_mat_gather_batch(self.h, args, gs, len(items)))

    def gather_finish(self, gather: "Gather", ll_src_ptr: int, dense_out_ptr: int, count: int):
        self._check(self.lib.b200_gather_finish(self.h, C.byref(gather), C.c_void_p(ll_src_ptr), C.c_void_p(dense_out_ptr), count))

    def make_args(self, w: QTensor, x_ptr: int, n: int, dst_ptr: int, m: int | None = None) -> MulMatArgs:
        a = MulMatArgs()
        a.type = w.type
        a.src0_dev = w.ptr
        a.src0_nblocks_total = w.nblocks
        a.ne00, a.ne01, a.ne02, a.ne03 = w.k, (w.m if m is None else m), w.ne02, w.ne03
        a.src1_dev = x_ptr
        a.ne11, a.ne12, a.ne13 = n, 1, 1
        a.nb11 = w.k * 4
        a.nb12 = a.nb13 = a.nb11 * n
        a.dst_dev = dst_ptr
        return a

    def mul_mat_fused(self, w: QTensor, x_ptr: int, n: int, dst_ptr: int, bias_ptr: int = 0, residual_ptr: int = 0, act: int = EPI_NONE, residual2_ptr: int = 0):
        """decode mul_mat with bias / GELU / residual folded into the GEMV epilogue (b200_mul_mat_fused)"""
        a = self.make_args(w, x_ptr, n, dst_ptr)
        e = Epilogue()
        e.bias_dev, e.residual_dev, e.act, e.residual2_dev = bias_ptr or None, residual_ptr or None, act, residual2_ptr or None
        self._check(self.lib.b200_mul_mat_fused(self.h, C.byref(a), C.byref(e)))

    def mul_mat_batch(self, args_list):
        """independent mul_mats in one call (b200_mul_mat_batch); same-input decode entries share a launch"""
        arr = (MulMatArgs * len(args_list))(*args_list)
        self._check(self.lib.b200_mul_mat_batch(self.h, arr, len(args_list)))

    # -- decode plans (b200_plan_*): a dependent sequence of decode mul_mats as one persistent launch
    def plan_arena_bytes(self, args_list, split: "PlanSplit | None" = None) -> int:
        arr = (MulMatArgs * len(args_list))(*args_list)
        return int(self.lib.b200_plan_arena_bytes(arr, len(args_list), C.byref(split) if split is not None else None))

    def plan_create(self, args_list, split: "PlanSplit | None" = None) -> int:
        arr = (MulMatArgs * len(args_list))(*args_list)
        h = C.c_void_p()
        self._check(self.lib.b200_plan_create(self.h, arr, len(args_list), C.byref(split) if split is not None else None, C.byref(h)))
        return h.value

    def plan_launch(self, plan: int):
        self._check(self.lib.b200_plan_launch(self.h, C.c_void_p(plan)))

    def plan_destroy(self, plan: int):
        self.lib.b200_plan_destroy(C.c_void_p(plan))

    def plan_trace(self, plan: int) -> np.ndarray:
        """[nops + 1, grid, 4] ns stamps of the last launch (needs set_option('plan_trace', 1) before plan_create); the last row holds
        per-CTA totals: producer blocked on a full ring, a consumer warp blocked on an empty ring, time in quantization phases"""
        nops, grid = C.c_int(), C.c_int()
        self.lib.b200_plan_trace(self.h, C.c_void_p(plan), None, 0, C.byref(nops), C.byref(grid))
        out = np.zeros((nops.value + 1, grid.value, 4), dtype=np.uint64)   # last row: per-CTA totals
        self._check(self.lib.b200_plan_trace(self.h, C.c_void_p(plan), _ptr(out), out.size, C.byref(nops), C.byref(grid)))
        return out

    def mul_mat(self, w: QTensor, x: np.ndarray, flags: int = 0) -> np.ndarray:
        """x: [ne13, ne12, n, k] (or [n, k]) float32 host.  Returns dst [ne13, ne12, n, m] float32 (ggml dst[m,n,ne12,ne13])."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        shape = x.shape
        if x.ndim == 2:
            x = x[None, None]
        ne13, ne12, n, k = x.shape
        assert k == w.k
        xd = self.to_device(x)
        out = self.alloc(ne13 * ne12 * n * w.m * 4)
        try:
            self.mul_mat_device(w, xd.ptr, n, out.ptr, ne12, ne13, flags=flags)
            self.synchronize()
            y = out.download(np.float32, ne13 * ne12 * n * w.m).reshape(ne13, ne12, n, w.m)
            return y[0, 0] if len(shape) == 2 else y
        finally:
            xd.free(); out.free()

    def mul_mat_host(self, w: QTensor, x: np.ndarray) -> np.ndarray:
        """End-to-end convenience call with HOST buffers (upload + mul_mat + download inside the C ABI)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, k = x.shape
        y = np.empty((n, w.m), dtype=np.float32)
        self._check(self.lib.b200_mul_mat_host(self.h, w.type, C.c_void_p(w.ptr), k, w.m, _ptr(x), n, _ptr(y)))
        return y

    def block_dots(self, w: QTensor, x: np.ndarray, path: int = 0) -> np.ndarray:
        """Per-block int32 partial sums [n, m, k/32] exactly as the kernels form them (parity instrumentation)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, k = x.shape
        nb = k // QK
        xd = self.to_device(x)
        out = self.alloc(n * w.m * nb * 4)
        try:
            self._check(self.lib.b200_memset(self.h, C.c_void_p(out.ptr), 0x7f, n * w.m * nb * 4))
            self._check(self.lib.b200_block_dots(self.h, w.type, C.c_void_p(w.ptr), k, w.m, C.c_void_p(xd.ptr), n, C.c_void_p(out.ptr), path))
            self.synchronize()
            return out.download(np.int32, n * w.m * nb).reshape(n, w.m, nb)
        finally:
            xd.free(); out.free()


# ---- synthetic weights made directly in the wire format (no CPU quantizer involved) ---------------------

def random_wire_weights(qtype: int, k: int, m: int, seed: int = 1234, scale: float | None = None) -> np.ndarray:
    """Random-init Q4_0/Q8_0 rows in wire format: uniform quants, fp16 block scales around `scale`
    (default: weights ~ unit-variance rows / sqrt(k), so chains of mul_mats stay O(1))."""
    rng = np.random.default_rng(seed)
    nb = k // QK
    wb = WIRE_BYTES[qtype]
    out = np.empty((m * nb, wb), dtype=np.uint8)
    qstd = 4.32 if qtype == TYPE_Q4_0 else 73.3
    base = (1.0 / (qstd * np.sqrt(k))) if scale is None else scale
    d = (base * rng.uniform(0.8, 1.2, size=m * nb)).astype(np.float16)
    out[:, 0:2] = d.view(np.uint8).reshape(-1, 2)
    if qtype == TYPE_Q4_0:
        # nibbles 1..15 -> quants -7..7: zero-mean, so a chain of mul_mats keeps O(1) activations
        out[:, 2:] = rng.integers(1, 16, size=(m * nb, 16), dtype=np.uint8) | (rng.integers(1, 16, size=(m * nb, 16), dtype=np.uint8) << 4)
    else:
        out[:, 2:] = rng.integers(-127, 128, size=(m * nb, 32), dtype=np.int8).view(np.uint8)
    return out.reshape(m, nb * wb)
