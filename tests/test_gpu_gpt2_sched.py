"""BASELINE.json configs[2]: GPT-2 117M with Q4_0 / Q8_0 matrices (random-init), a 128-token prompt and decode steps, on the
reference's own scheduler (ggml_backend_sched over {B200, CPU}, quantized MUL_MAT nodes pinned to the B200 backend, the glue ops on
the CPU backend) against the same model computed entirely by the reference CPU backend: oracle/gpt2_sched_harness.c, built by
oracle/Makefile into oracle/_ref/gpt2-sched-harness.  The tied wte lives repacked in a B200 buffer and reaches GET_ROWS on the
CPU through the scheduler's tensor copy (get_tensor's exact un-repack).  A third arm computes the WHOLE graph on the B200 backend
(compute tensors from ggml_gallocr on its buffer type, one ggml_backend_graph_compute per step: what examples/gpt-2/main-backend.cpp does).
Logits within test-backend-ops' NMSE bar 5e-4."""
import json
import subprocess

import pytest

from conftest import ROOT, MUL_MAT_NMSE_TOL

pytestmark = pytest.mark.gpu
HARNESS = ROOT / "oracle" / "_ref" / "gpt2-sched-harness"


@pytest.mark.parametrize("qname,parallel", [("q4_0", 0), ("q8_0", 0), ("q4_0", 1)])
def test_gpt2_117m_prompt_and_decode_on_sched_b200_plus_cpu(qname, parallel):
    """parallel = 1: ggml_backend_sched_new(..., parallel = true) -- the scheduler's pipelined mode, which goes through this backend's
    event_new / event_record / event_wait / event_synchronize and set_tensor_async entries (src/ggml-backend.c:1640-1712)"""
    assert HARNESS.exists(), f"{HARNESS} must be prebuilt (make -C oracle dropin) and travel with the snapshot"
    p = subprocess.run([str(HARNESS), qname, "128", "3", "8", "1", str(parallel)], capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, f"no result line (rc {p.returncode}): {p.stdout[-400:]} {p.stderr[-800:]}"
    r = json.loads(lines[-1])
    assert "error" not in r, r
    assert r["quantized_mul_mat_nodes_on_b200"] == 12 * 4 + 1          # qkv, proj, fc, out per block + the tied head
    steps = r["steps"]
    assert [s["n"] for s in steps] == [128, 1, 1, 1] and [s["n_past"] for s in steps] == [0, 128, 129, 130]
    for s in steps:
        assert s["finite"] and s["logits_nmse_vs_cpu"] <= MUL_MAT_NMSE_TOL, s
        assert s["b200_whole_graph_logits_nmse_vs_cpu"] <= MUL_MAT_NMSE_TOL, s
        assert 0 < s["b200_launches"] < s["graph_nodes"], s          # views cost nothing, in-place neighbours share a kernel
    assert r["b200_fused_nodes_total"] > 0
    assert r["ok"] and p.returncode == 0
