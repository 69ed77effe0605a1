"""Pins tests/rope_restatement.py to the reference itself: the unmodified reference core (oracle/_ref/libggml.so, built by oracle/Makefile from
/root/reference where it lies) computes GGML_OP_ROPE on its CPU path through its public API (ggml_rope_custom / ggml_rope_xpos_inplace +
ggml_graph_compute_with_ctx, src/ggml.c:13775, :13953), and the numpy restatement has to agree on the same inputs."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT, nmse
from rope_restatement import rope_numpy

LIB = ROOT / "oracle" / "_ref" / "libggml.so"


class InitParams(C.Structure):
    _fields_ = [("mem_size", C.c_size_t), ("mem_buffer", C.c_void_p), ("no_alloc", C.c_bool)]


@pytest.fixture(scope="module")
def ggml():
    if not LIB.exists():
        pytest.skip("oracle/_ref/libggml.so not built (make -C oracle dropin)")
    g = C.CDLL(str(LIB))
    g.ggml_init.restype = C.c_void_p
    g.ggml_init.argtypes = [InitParams]
    g.ggml_new_tensor_4d.restype = C.c_void_p
    g.ggml_new_tensor_4d.argtypes = [C.c_void_p, C.c_int] + [C.c_int64] * 4
    g.ggml_new_tensor_1d.restype = C.c_void_p
    g.ggml_new_tensor_1d.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    g.ggml_rope_custom.restype = C.c_void_p
    g.ggml_rope_custom.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_float] * 6
    g.ggml_rope_xpos_inplace.restype = C.c_void_p
    g.ggml_rope_xpos_inplace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_bool]
    g.ggml_new_graph.restype = C.c_void_p
    g.ggml_new_graph.argtypes = [C.c_void_p]
    g.ggml_build_forward_expand.argtypes = [C.c_void_p, C.c_void_p]
    g.ggml_graph_compute_with_ctx.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    g.ggml_get_data.restype = C.c_void_p
    g.ggml_get_data.argtypes = [C.c_void_p]
    g.ggml_free.argtypes = [C.c_void_p]
    return g


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("ne0,heads,n_dims,mode,yarn,xpos", [(64, 16, 64, 0, False, False), (128, 5, 128, 0, True, False), (64, 7, 64, 2, False, False),
                                                             (80, 4, 20, 2, True, False), (128, 3, 128, 0, False, True)])
def test_rope_restatement_matches_reference_cpu(ggml, dtype, ne0, heads, n_dims, mode, yarn, xpos):
    if xpos and dtype == np.float16:
        pytest.skip("the reference applies the xPos factor on F32 only")
    rng = np.random.default_rng(ne0 + heads + mode)
    B, T = 2, 9
    x = rng.uniform(-1, 1, (B, T, heads, ne0)).astype(dtype)
    pos = rng.integers(0, 512, T).astype(np.int32)
    kw = dict(n_dims=n_dims, mode=mode, n_orig_ctx=256 if yarn else 0, freq_base=10000.0, freq_scale=0.5 if yarn else 1.0, ext_factor=0.7 if yarn else 0.0,
              attn_factor=1.1 if yarn else 1.0, beta_fast=32.0 if yarn else 0.0, beta_slow=1.0 if yarn else 0.0)
    if xpos:
        kw.update(xpos_base=512.0, xpos_down=True)
    want = rope_numpy(x, pos, **kw)
    ctx = ggml.ggml_init(InitParams(64 << 20, None, False))
    try:
        a = ggml.ggml_new_tensor_4d(ctx, 0 if dtype == np.float32 else 1, ne0, heads, T, B)
        p = ggml.ggml_new_tensor_1d(ctx, 26, T)                                  # GGML_TYPE_I32
        C.memmove(ggml.ggml_get_data(a), x.ctypes.data, x.nbytes)
        C.memmove(ggml.ggml_get_data(p), pos.ctypes.data, pos.nbytes)
        if xpos:
            out = ggml.ggml_rope_xpos_inplace(ctx, a, p, n_dims, 512.0, True)
        else:
            out = ggml.ggml_rope_custom(ctx, a, p, n_dims, mode, 512, kw["n_orig_ctx"], kw["freq_base"], kw["freq_scale"], kw["ext_factor"], kw["attn_factor"],
                                        kw["beta_fast"], kw["beta_slow"])
        gf = ggml.ggml_new_graph(ctx)
        ggml.ggml_build_forward_expand(gf, out)
        assert ggml.ggml_graph_compute_with_ctx(ctx, gf, 1) == 0
        got = np.empty_like(x)
        C.memmove(got.ctypes.data, ggml.ggml_get_data(out), x.nbytes)
    finally:
        ggml.ggml_free(ctx)
    assert nmse(got.astype(np.float64), want.astype(np.float64)) <= (1e-10 if dtype == np.float32 else 1e-7)
    assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= (5e-6 if dtype == np.float32 else 1e-3)
