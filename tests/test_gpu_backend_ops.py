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


def test_backend_ops_all_ops_do_not_crash():
    """Every other op must be declined through supports_op (printed 'not supported'), never attempted."""
    exe = REF / "test-backend-ops"
    rc, out = run([str(exe), "test", "-b", "B2000"])
    assert rc == 0, out[-4000:]
