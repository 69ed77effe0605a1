"""The reference's OWN conformance harness (tests/test-backend-ops.cpp, tests/test-backend-buffer.cpp), compiled
unmodified in the build container into oracle/_ref/ against the reference core + this backend, run on the B200.
It discovers the backend through the registry (ggml_backend_cuda_reg_devices alias) and compares every
MUL_MAT case with the reference CPU backend in-process (NMSE <= 5e-4)."""
import re
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = ROOT / "oracle" / "_ref"


def run(args):
    p = subprocess.run(args, cwd=REF, capture_output=True, text=True, timeout=900)
    return p.returncode, p.stdout + p.stderr


def test_backend_buffer_smoke():
    exe = REF / "test-backend-buffer"
    assert exe.exists(), "oracle/_ref/test-backend-buffer must be prebuilt (make -C oracle dropin) and travel with the snapshot"
    rc, out = run([str(exe)])
    assert rc == 0, out
    assert "B2000" in out, out


def test_backend_ops_mul_mat():
    exe = REF / "test-backend-ops"
    assert exe.exists(), "oracle/_ref/test-backend-ops must be prebuilt (make -C oracle dropin)"
    rc, out = run([str(exe), "test", "-o", "MUL_MAT", "-b", "B2000"])
    assert rc == 0, out[-4000:]
    clean = re.sub(r"\x1b\[[0-9;]*m", "", out)
    q = [l for l in clean.splitlines() if "MUL_MAT(type_a=q4_0,type_b=f32" in l or "MUL_MAT(type_a=q8_0,type_b=f32" in l]
    assert len(q) >= 15, clean[-4000:]
    assert all(l.rstrip().endswith("OK") for l in q), "\n".join(q)
    assert re.search(r"\d+/\d+ tests passed", clean) and "FAIL" not in clean, clean[-2000:]


GLUE = ["GET_ROWS(type=f32", "GET_ROWS(type=f16", "GET_ROWS(type=q4_0", "GET_ROWS(type=q8_0", "ADD(type=f32", "MUL(type=f32", "DIV(type=f32", "GELU(type=f32",
        "GELU_QUICK(type=f32", "SILU(type=f32", "RELU(type=f32", "TANH(type=f32", "NORM(type=f32", "RMS_NORM(type=f32", "SCALE(type=f32", "DIAG_MASK_INF(type=f32",
        "SOFT_MAX(type=f32", "CPY(type_src=f32,type_dst=f32", "CPY(type_src=f32,type_dst=f16", "CPY(type_src=f16,type_dst=f32", "DUP(type=f32", "DUP(type=i16",
        "CONT(type=f32", "MUL_MAT(type_a=f32,type_b=f32", "MUL_MAT(type_a=f16,type_b=f32", "MUL_MAT_ID(type_a=q4_0,type_b=f32", "MUL_MAT_ID(type_a=q8_0,type_b=f32",
        # the sibling 32-element formats that share the Q8_0 activation path
        "MUL_MAT(type_a=q5_0,type_b=f32", "MUL_MAT(type_a=iq4_nl,type_b=f32", "GET_ROWS(type=q5_0", "GET_ROWS(type=iq4_nl", "MUL_MAT_ID(type_a=q5_0,type_b=f32",
        "MUL_MAT_ID(type_a=iq4_nl,type_b=f32",
        # what a GPT-J graph adds
        "ROPE(type=f32", "ROPE(type=f16", "REPEAT(type=f32", "REPEAT(type=i32", "REPEAT(type=i16"]


def test_backend_ops_whole_suite_glue_ops_green_rest_declined():
    """The reference's whole conformance suite against this backend: every case of the operators either side of the path (SURVEY.md
    8(f)-1: what a GPT-2 / GPT-J graph computes between its quantized mul_mats) must RUN and match the CPU backend within the
    harness' own per-op bounds; everything else must be declined through supports_op (printed 'not supported'), never attempted."""
    exe = REF / "test-backend-ops"
    rc, out = run([str(exe), "test", "-b", "B2000"])
    clean = re.sub(r"\x1b\[[0-9;]*m", "", out)
    assert rc == 0 and "FAIL" not in clean, clean[-4000:]
    lines = clean.splitlines()
    for prefix in GLUE:
        mine = [l for l in lines if l.strip().startswith(prefix)]
        assert mine, f"no test case starts with {prefix}"
        assert all(l.rstrip().endswith("OK") for l in mine), "\n".join(l for l in mine if not l.rstrip().endswith("OK"))
    ran = sum(1 for l in lines if l.rstrip().endswith("OK"))
    assert ran >= 300, ran
