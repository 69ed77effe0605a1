#!/usr/bin/env python3
"""A/B of the decode-plan options on the GPT-J-6B Q4_0 decode graph (bench.py's workload), one process, weights uploaded once.
Every variant = set options, plan_create, 3 warm-ups, `reps` timed launches (CUDA events); the logits of every variant are
compared bit for bit with the first one.  `trace` variants add the device-side timeline of one launch.
Usage on the GPU box:  python tools/ab_plan.py [sweep|timeline]"""
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
dag = bench.gptj_dag()                      # order: v, fc_in, q, k, o, fc_out
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
DEFAULTS = {"plan_pub_min_k": 4096, "plan_pub_dist": 2, "plan_l2_window": 8, "plan_evict_first": 1, "plan_slots": 0, "plan_trace": 0}
ORDERS = {"v,fc_in,q,k,o,fc_out": [0, 1, 2, 3, 4, 5], "fc_in,v,q,k,o,fc_out": [1, 0, 2, 3, 4, 5], "v,q,fc_in,k,o,fc_out": [0, 2, 1, 3, 4, 5]}


def reorder(perm):
    """the same graph with every block's nodes permuted: [(weight index, src position in the new order)]"""
    order = []
    for b in range(0, n - 1, 6):
        order += [b + j for j in perm]
    order.append(n - 1)
    pos = {old: new for new, old in enumerate(order)}
    return [(old, -1 if dag[old][3] < 0 else pos[dag[old][3]]) for old in order]

BYTES = sum(bench.algorithmic_bytes(m, k, 1, 18) for _, m, k, _ in dag)


def run(opts, trace=False, reps=60, timeline=False, order="v,fc_in,q,k,o,fc_out"):
    nodes = reorder(ORDERS[order])
    cfg = dict(DEFAULTS)
    cfg.update(opts)
    cfg["plan_trace"] = 1 if trace else 0
    for key, v in cfg.items():
        ctx.set_option(key, v)
    label = ",".join(f"{k[5:]}={v}" for k, v in cfg.items() if k != "plan_trace")
    lens = [((weights[w].m + 15) // 16) * 16 for w, _ in nodes]
    at = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = torch.zeros(int(at[-1]), dtype=torch.float32, device=dev)
    args = [ctx.make_args(weights[w], x4.data_ptr() if src < 0 else out.data_ptr() + int(at[src]) * 4, 1, out.data_ptr() + int(at[i]) * 4)
            for i, (w, src) in enumerate(nodes)]
    plan = ctx.plan_create(args)
    for _ in range(3):
        ctx.plan_launch(plan)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.plan_launch(plan)
    e1.record(stream)
    ctx.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    extra = ""
    if trace:
        tr_all = ctx.plan_trace(plan).astype(np.int64)
        tot = tr_all[-1]
        extra = (f"  per CTA us: producer blocked {tot[:, 0].mean() / 1e3:6.1f}  consumer(w2) blocked {tot[:, 1].mean() / 1e3:6.1f}  "
                 f"new-input phases {tot[:, 2].mean() / 1e3:6.1f}")
        if timeline:
            tr = tr_all[:-1]
            live = tr[tr > 0]
            t0 = live.min()
            print(f"   timeline of the last launch, ns since the first stamp, min..max over CTAs (whole launch {(live.max() - t0) / 1e3:.1f} us)")
            print("    op name      " + " ".join(f"{c:>19s}" for c in ("src1 complete", "quantized", "first weights", "last row")))
            for i in list(range(13)) + [n - 2, n - 1]:
                row = []
                for sidx in range(4):
                    v = tr[i, :, sidx]
                    v = v[v > 0] - t0
                    row.append(f"{v.min():8d}..{v.max():8d}" if v.size else " " * 18)
                print(f"    {i:3d} {dag[nodes[i][0]][0]:8s} " + "  ".join(row))
    last = out[int(at[-2]):int(at[-2]) + 50400].cpu().numpy()
    ctx.plan_destroy(plan)
    label = ("" if order.startswith("v,fc_in") else order + " ") + label
    print(f"{label:80s} {us:8.1f} us/token  {1e6 / us:7.1f} tok/s  {BYTES / us / 1e3:7.1f} GB/s{extra}", flush=True)
    return last


which = sys.argv[1] if len(sys.argv) > 1 else "sweep"
base = run({})
if which == "sweep":
    for v in ({"plan_evict_first": 0}, {"plan_l2_window": 0}, {"plan_l2_window": 4}, {"plan_l2_window": 6}, {"plan_l2_window": 12}, {"plan_l2_window": 16},
              {"plan_l2_window": 24}, {"plan_l2_window": 16, "plan_evict_first": 0}, {"plan_pub_min_k": 8192}, {"plan_pub_min_k": 8192, "plan_l2_window": 16},
              {"plan_pub_min_k": 0}, {"plan_pub_dist": 1}):
        got = run(v)
        assert np.array_equal(base.view(np.uint32), got.view(np.uint32)), f"{v}: plan output differs"
    for order in list(ORDERS)[1:]:
        for v in ({}, {"plan_l2_window": 16}, {"plan_pub_min_k": 8192}):
            got = run(v, order=order)
            assert np.array_equal(base.view(np.uint32), got.view(np.uint32)), f"{order} {v}: plan output differs"
run({}, trace=True, reps=5, timeline=True)
run({}, trace=True, reps=5, timeline=True, order="fc_in,v,q,k,o,fc_out")
print("bitwise equal across variants: True")
