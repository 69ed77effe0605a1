"""The C-ABI library loads without a GPU and exports exactly what include/ggml_b200.h declares; the ggml
backend library exports the native entry points and the ggml-cuda.h alias set."""
import re
import subprocess

import pytest

from conftest import ROOT, PKG

HEADER = ROOT / "include" / "ggml_b200.h"


def header_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"B200_API[^;(]*?\b(b200_\w+)\s*\(", text)))


def exported(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(lib)], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_header_declares_symbols():
    syms = header_symbols()
    assert len(syms) >= 25 and "b200_mul_mat" in syms and "b200_set_quantized" in syms


def test_library_exports_every_declared_symbol(qmm):
    assert qmm.LIB_PATH.exists(), "build the extension first (__graft_entry__.build())"
    exp = exported(qmm.LIB_PATH)
    missing = [s for s in header_symbols() if s not in exp]
    assert not missing, missing


def test_binding_covers_header(qmm):
    assert sorted(qmm.declared_symbols()) == header_symbols()
    lib = qmm.load_library()   # dlopen + resolve every symbol; works without a GPU
    assert lib.b200_device_count() >= 0


def test_no_cpu_fallback_without_device(qmm):
    """Without a usable sm_100 device context creation must fail loudly, never fall back."""
    if qmm.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(qmm.B200Error):
        qmm.Context(0)


def test_backend_library_exports():
    lib = PKG / "lib" / "libggml-b200-backend.so"
    if not lib.exists():
        pytest.skip("backend library needs the reference headers at build time")
    exp = exported(lib)
    for s in ["ggml_backend_b200_init", "ggml_backend_is_b200", "ggml_backend_b200_buffer_type",
              "ggml_backend_b200_host_buffer_type", "ggml_backend_b200_get_device_count", "ggml_backend_b200_reg_devices",
              # src/ggml-cuda.h:19-39 alias set
              "ggml_backend_cuda_init", "ggml_backend_is_cuda", "ggml_backend_cuda_buffer_type",
              "ggml_backend_cuda_host_buffer_type", "ggml_backend_cuda_get_device_count",
              "ggml_backend_cuda_get_device_description", "ggml_backend_cuda_get_device_memory",
              "ggml_backend_cuda_register_host_buffer", "ggml_backend_cuda_unregister_host_buffer",
              "ggml_backend_cuda_reg_devices"]:
        assert s in exp, s


def test_backend_library_exports_every_symbol_of_its_header():
    """host/ggml-b200.h (what an application includes next to ggml-backend.h): every declared entry point is exported"""
    lib = PKG / "lib" / "libggml-b200-backend.so"
    if not lib.exists():
        pytest.skip("backend library needs the reference headers at build time")
    text = re.sub(r"/\*.*?\*/", "", (PKG / "host" / "ggml-b200.h").read_text(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(ggml_backend_b200_\w+)\s*\(", text)))
    assert len(declared) >= 12 and "ggml_backend_b200_graph_plan_kernels" in declared and "ggml_backend_b200_split_buffer_type" in declared
    exp = exported(lib)
    missing = [d for d in declared if d not in exp]
    assert not missing, missing


def test_product_does_not_reference_oracle():
    """No file of the product tree may include, link or call anything under oracle/."""
    for p in list((PKG / "csrc").glob("*")) + list((PKG / "host").glob("*")) + [PKG / "qmm.py", PKG / "Makefile", HEADER]:
        txt = p.read_text()
        assert "qmm_oracle" not in txt and "oracle/" not in txt and "libref_shim" not in txt, p
