"""SURVEY.md 8(f)-1, the part GPT-J adds (ROPE, REPEAT, an F16 KV cache written through strided copies and multiplied through permuted views):
the north-star model's WHOLE graph, built with the reference's public API the way examples/gpt-j/main.cpp:440-586 computes it, on the B200
backend against the reference CPU backend -- oracle/gptj_harness.c, built by oracle/Makefile into oracle/_ref/gptj-harness.  A scaled-down GPT-J
here (the 6B run is a profile: profiles/r02_gptj_6b_whole_graph_q4_0.json); logits of the last token within test-backend-ops' NMSE bar 5e-4, and the
recorded graph plan bitwise equal to the node-by-node compute."""
import json
import subprocess

import pytest

from conftest import ROOT, MUL_MAT_NMSE_TOL

pytestmark = pytest.mark.gpu
HARNESS = ROOT / "oracle" / "_ref" / "gptj-harness"


@pytest.mark.parametrize("qname,fuse,taps", [("q4_0", 1, 0), ("q8_0", 1, 0), ("q4_0", 0, 0), ("q4_0", 1, 1)])
def test_gptj_whole_graph_prompt_and_decode(qname, fuse, taps):
    """taps = 1: two extra outputs read intermediates the fusions would skip (a LayerNorm's product before its bias, an MLP's pre-activation): the backend
    has to see the second reader, leave those groups unfused and produce the taps the CPU produces"""
    assert HARNESS.exists(), f"{HARNESS} must be prebuilt (make -C oracle dropin) and travel with the snapshot"
    #                                   n_layer n_embd n_head n_rot n_vocab n_ctx n_prompt n_decode threads
    args = [str(HARNESS), qname, "3", "512", "8", "32", "2000", "128", "33", "3", "8", str(fuse), "1", str(taps)]
    p = subprocess.run(args, capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, f"no result line (rc {p.returncode}): {p.stdout[-400:]} {p.stderr[-800:]}"
    r = json.loads(lines[-1])
    assert "error" not in r, r
    steps = r["steps"]
    assert [s["n"] for s in steps] == [33, 1, 1, 1] and [s["n_past"] for s in steps] == [0, 33, 34, 35]
    for s in steps:
        assert s["finite"] and s["logits_nmse_vs_cpu"] <= MUL_MAT_NMSE_TOL and s["taps_nmse_vs_cpu"] <= 1e-6, s
        assert 0 < s["b200_launches"] < s["graph_nodes"], s
        if s["n"] == 1:
            assert s["graph_plan_kernels"] > 0 and s["graph_plan_equals_node_by_node"], s
    assert (r["b200_fused_nodes_total"] > 0) == bool(fuse)
    assert r["ok"] and p.returncode == 0
