"""a GGUF checkpoint -- written and parsed by the reference's own gguf_* API -- goes through a pinned host staging buffer
(ggml_backend_cuda_host_buffer_type) into the backend's repacked device layout: oracle/gguf_load_harness.c, built into oracle/_ref/gguf-load-harness.
Every tensor read back with ggml_backend_tensor_get equals the file's bytes; every matrix (Q4_0, Q8_0, one Q5_0, one IQ4_NL) multiplied on the B200
backend matches the reference CPU backend computing on the file's bytes."""
import json
import subprocess

import pytest

from conftest import ROOT, MUL_MAT_NMSE_TOL

pytestmark = pytest.mark.gpu
HARNESS = ROOT / "oracle" / "_ref" / "gguf-load-harness"


def test_gguf_file_into_repacked_device_layout(tmp_path):
    assert HARNESS.exists(), f"{HARNESS} must be prebuilt (make -C oracle dropin) and travel with the snapshot"
    p = subprocess.run([str(HARNESS), "1024", "2", str(tmp_path / "model.gguf")], capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, f"no result line (rc {p.returncode}): {p.stdout[-400:]} {p.stderr[-800:]}"
    r = json.loads(lines[-1])
    assert "error" not in r, r
    assert r["tensors"] == 14 and r["loaded_bytes"] == r["payload_bytes"]
    assert r["tensor_get_equals_file"] and r["mul_mats_checked"] == 12 and r["worst_nmse_vs_cpu"] <= MUL_MAT_NMSE_TOL
    assert r["ok"] and p.returncode == 0
