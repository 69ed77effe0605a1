#!/usr/bin/env python3
"""A/B of B200_PLAN_LL_RING (src1 vectors through the weight ring) on the GPT-J-6B Q4_0 decode graph, both node orders.
One process, the weights uploaded once; every variant = plan_create + 3 warm-ups + 60 timed launches (CUDA events), then
one traced launch for the per-CTA totals (ring misses = runs a warp had to re-fetch from L2 because the copy was early).
The last output vector of every variant is compared bit for bit with the first variant of the same order."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

qmm = bench.load_qmm()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = qmm.Context(0, stream=stream.cuda_stream)
dag = bench.gptj_dag()                      # default order: fc_in, v, q, k, o, fc_out
host_w, weights, keep = {}, [], []
for name, m, k, _ in dag:
    if (m, k) not in host_w:
        host_w[(m, k)] = qmm.random_wire_weights(2, k, m, seed=1234 + m + k)
    buf = torch.empty(m * (k // 32) * 18, dtype=torch.uint8, device=dev)
    keep.append(buf)
    t = qmm.QTensor(ctx, 2, k, m, ptr=buf.data_ptr())
    t.set(host_w[(m, k)])
    weights.append(t)
x4 = torch.rand(4096, device=dev) * 2 - 1
n = len(dag)
default_nodes = [(i, dag[i][3]) for i in range(n)]


def v_first(nodes):
    """the same graph with v moved in front of fc_in in every block: [(weight index, src position)]"""
    order = []
    for b in range(0, n - 1, 6):
        order += [b + 1, b + 0, b + 2, b + 3, b + 4, b + 5]
    order.append(n - 1)
    pos = {old: new for new, old in enumerate(order)}
    return [(old, -1 if nodes[old][1] < 0 else pos[nodes[old][1]]) for old in order]


KNOBS = ("B200_PLAN_NOSPLIT", "B200_PLAN_PUBQ", "B200_PLAN_LLQ", "B200_PLAN_LLQ_DIST", "B200_PLAN_LL_RING", "B200_PLAN_L2_SLOTS", "B200_PLAN_L2_AHEAD", "B200_PLAN_SLOTS", "B200_PLAN_TRACE")


def run(label, nodes, env=None, trace=False, reps=60, timeline=False):
    env = dict(env or {})
    env.setdefault("B200_PLAN_LLQ", 0)          # (on by default on one GPU: every variant names it)
    if trace:
        env["B200_PLAN_TRACE"] = 1
    for key in KNOBS:
        if key in env:
            os.environ[key] = str(env[key])
        else:
            os.environ.pop(key, None)
    ring = ",".join(f"{k[10:]}={v}" for k, v in env.items() if k != "B200_PLAN_TRACE") or "shipped"
    lens = [((weights[w].m + 15) // 16) * 16 for w, _ in nodes]
    at = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = torch.zeros(int(at[-1]), dtype=torch.float32, device=dev)
    args = [ctx.make_args(weights[w], x4.data_ptr() if src < 0 else out.data_ptr() + int(at[src]) * 4, 1, out.data_ptr() + int(at[i]) * 4)
            for i, (w, src) in enumerate(nodes)]
    plan = ctx.plan_create(args)
    for _ in range(3):
        ctx.plan_launch(plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.plan_launch(plan)
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    extra = ""
    if trace:
        tot = ctx.plan_trace(plan).astype(np.int64)[-1]
        extra = (f"  per CTA us: producer blocked {tot[:, 0].mean() / 1e3:6.1f}  consumer(w2) blocked {tot[:, 1].mean() / 1e3:6.1f}  "
                 f"quantize phases {tot[:, 2].mean() / 1e3:6.1f}  ring misses (w2, all launches) {tot[:, 3].sum()}")
        if timeline:
            tr = ctx.plan_trace(plan).astype(np.int64)[:-1]
            live = tr[tr > 0]
            t0 = live.min()
            print(f"   timeline of the last launch, ns since the first stamp, min..max over CTAs (whole launch {(live.max() - t0) / 1e3:.1f} us)")
            print("    op name      " + " ".join(f"{c:>19s}" for c in ("src1 complete", "quantized", "first weights", "last row")))
            for i in list(range(13)) + [len(nodes) - 2, len(nodes) - 1]:
                row = []
                for sidx in range(4):
                    v = tr[i, :, sidx]
                    v = v[v > 0] - t0
                    row.append(f"{v.min():8d}..{v.max():8d}" if v.size else " " * 18)
                print(f"    {i:3d} {dag[nodes[i][0]][0]:8s} " + "  ".join(row))
    last = out[int(at[-2]):int(at[-2]) + 50400].cpu().numpy()
    ctx.plan_destroy(plan)
    print(f"{label:30s} {ring:14s} {us:8.1f} us/token  {1e6 / us:7.1f} tok/s{extra}", flush=True)
    return last


which = sys.argv[1] if len(sys.argv) > 1 else "l2"
orders = (("fc_in,v,q,k,o,fc_out", default_nodes), ("v,fc_in,q,k,o,fc_out", v_first(default_nodes)))
if which == "ring":
    for oname, nodes in orders:
        base = run(oname, nodes)
        for ring in (6, 12, 20):
            got = run(oname, nodes, {"B200_PLAN_LL_RING": ring})
            assert np.array_equal(base.view(np.uint32), got.view(np.uint32)), "ring-fed plan differs"
        run(oname + " [traced]", nodes, trace=True, reps=5)
        run(oname + " [traced]", nodes, {"B200_PLAN_LL_RING": 12}, trace=True, reps=5)
elif which == "llq":
    oname, nodes = orders[0]
    base = run(oname, nodes)
    for env in ({"B200_PLAN_LLQ": 8192}, {"B200_PLAN_LLQ": 4096}, {"B200_PLAN_LLQ": 4096, "B200_PLAN_LLQ_DIST": 1}):
        got = run(oname, nodes, env)
        assert np.array_equal(base.view(np.uint32), got.view(np.uint32)), f"{env}: plan output differs"
    run(oname + " [traced]", nodes, {"B200_PLAN_LLQ": 8192}, trace=True, reps=5, timeline=True)
elif which == "pubq":          # the experimental kernel modes: 8 = published planes + arrival counter + one bulk copy per CTA, 16 = no k-split teams
    oname, nodes = orders[0]
    base = run(oname, nodes)
    for env in ({"B200_PLAN_LLQ": 8192}, {"B200_PLAN_LLQ": 8192, "B200_PLAN_PUBQ": 1}, {"B200_PLAN_LLQ": 4096, "B200_PLAN_PUBQ": 1},
                {"B200_PLAN_LLQ": 4096, "B200_PLAN_PUBQ": 1, "B200_PLAN_LLQ_DIST": 1},
                {"B200_PLAN_NOSPLIT": 1}, {"B200_PLAN_LLQ": 8192, "B200_PLAN_NOSPLIT": 1}, {"B200_PLAN_LLQ": 8192, "B200_PLAN_PUBQ": 1, "B200_PLAN_NOSPLIT": 1},
                {"B200_PLAN_LLQ": 4096, "B200_PLAN_PUBQ": 1, "B200_PLAN_NOSPLIT": 1}):
        got = run(oname, nodes, env)
        assert np.array_equal(base.view(np.uint32), got.view(np.uint32)), f"{env}: plan output differs"
    run(oname + " [traced]", nodes, {"B200_PLAN_LLQ": 8192, "B200_PLAN_PUBQ": 1}, trace=True, reps=5, timeline=True)
    run(oname + " [traced]", nodes, {"B200_PLAN_LLQ": 4096, "B200_PLAN_PUBQ": 1}, trace=True, reps=5, timeline=True)
else:
    oname, nodes = orders[0]
    base = run(oname, nodes)
    for env in ({"B200_PLAN_L2_SLOTS": 4}, {"B200_PLAN_L2_SLOTS": 8}, {"B200_PLAN_L2_SLOTS": 16}, {"B200_PLAN_L2_SLOTS": 32},
                {"B200_PLAN_L2_SLOTS": 64}, {"B200_PLAN_L2_AHEAD": 1}, {"B200_PLAN_L2_SLOTS": 8, "B200_PLAN_SLOTS": 6}):
        got = run(oname, nodes, env)
        assert np.array_equal(base.view(np.uint32), got.view(np.uint32)), f"{env}: plan output differs"
    run(oname + " [traced]", nodes, trace=True, reps=5, timeline=True)
    run(oname + " [traced]", nodes, {"B200_PLAN_L2_SLOTS": 16}, trace=True, reps=5, timeline=True)
print("bitwise equal across variants: True")
