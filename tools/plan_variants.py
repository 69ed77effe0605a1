#!/usr/bin/env python3
"""Where does the decode plan lose time?  Same GPT-J-6B Q4_0 weights, different dependency structures:
  deps      the bench graph (28 x [q,k,v,fc_in <- x; o <- v; fc_out <- fc_in] + lm_head)
  nodeps    every op reads an outside vector (pure streaming: what the ring/consumers can do without any hand-off)
  chain4096 only the k = 4096 ops, each reading the previous one (worst-case dependency density)
Prints ms/token and effective TB/s per variant (B200_PLAN_SLOTS=n narrows the ring)."""
import sys
import time
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
dag = bench.gptj_dag()
host_w = {}
weights = []
keep = []
for name, m, k, _ in dag:
    if (m, k) not in host_w:
        host_w[(m, k)] = qmm.random_wire_weights(2, k, m, seed=1234 + m + k)
    buf = torch.empty(m * (k // 32) * 18, dtype=torch.uint8, device=dev)
    keep.append(buf)
    t = qmm.QTensor(ctx, 2, k, m, ptr=buf.data_ptr())
    t.set(host_w[(m, k)])
    weights.append(t)
x4 = torch.rand(4096, device=dev) * 2 - 1
x16 = torch.rand(16384, device=dev) * 2 - 1


import os
os.environ["B200_PLAN_TRACE"] = "1"


def run(label, nodes, slots=None, l2=None, rows=None):
    """nodes: [(weight index, src)] src = index into nodes or -1"""
    for key, val in (("B200_PLAN_SLOTS", slots), ("B200_PLAN_L2_AHEAD", l2), ("B200_PLAN_SLOT_ROWS", rows)):
        if val is not None:
            os.environ[key] = str(val)
        else:
            os.environ.pop(key, None)
    lens = [((weights[w].m + 15) // 16) * 16 for w, _ in nodes]
    at = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = torch.zeros(int(at[-1]), dtype=torch.float32, device=dev)
    args = []
    for i, (w, src) in enumerate(nodes):
        t = weights[w]
        if src < 0:
            sp = (x4 if t.k == 4096 else x16).data_ptr()
        else:
            sp = out.data_ptr() + int(at[src]) * 4
        args.append(ctx.make_args(t, sp, 1, out.data_ptr() + int(at[i]) * 4))
    plan = ctx.plan_create(args)
    for _ in range(3):
        ctx.plan_launch(plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record(stream)
    for _ in range(reps):
        ctx.plan_launch(plan)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = sum(weights[w].nbytes for w, _ in nodes)
    tot = ctx.plan_trace(plan).astype(np.int64)[-1]
    print(f"{label:28s} {len(nodes):4d} ops  {ms * 1e3:8.1f} us/launch  {by / ms / 1e9:6.2f} TB/s  ({by / 1e6:.0f} MB)   "
          f"us mean/max per CTA: producer blocked {tot[:, 0].mean() / 1e3:6.1f}/{tot[:, 0].max() / 1e3:6.1f}  consumer(w2) blocked "
          f"{tot[:, 1].mean() / 1e3:6.1f}/{tot[:, 1].max() / 1e3:6.1f}  quantize {tot[:, 2].mean() / 1e3:6.1f}/{tot[:, 2].max() / 1e3:6.1f}", flush=True)
    ctx.plan_destroy(plan)


# box sanity: plain device copy bandwidth (read + write), the way MEASURED_PEAKS.json is taken
a = torch.empty(1 << 29, dtype=torch.bfloat16, device=dev)
b = torch.empty_like(a)
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); b.copy_(a); e1.record(stream); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"copy bandwidth on this box: {2 * a.numel() * 2 / best / 1e6:.0f} GB/s", flush=True)
del a, b

n = len(dag)
run("deps (bench graph)", [(i, dag[i][3]) for i in range(n)])
run("deps, 6 ring slots", [(i, dag[i][3]) for i in range(n)], slots=6)
run("nodeps", [(i, -1) for i in range(n)])
k4 = [i for i in range(n) if dag[i][2] == 4096 and dag[i][1] == 4096]
run("k4096 m4096 nodeps", [(i, -1) for i in k4])
run("k4096 m4096 chain", [(i, j - 1) for j, i in enumerate(k4)])
big = [i for i in range(n) if dag[i][0] == "fc_in"]
run("fc_in only nodeps", [(i, -1) for i in big])
fo = [i for i in range(n) if dag[i][0] == "fc_out"]
run("fc_out only nodeps", [(i, -1) for i in fo])
# fc_in -> fc_out pairs chained (k = 16384 hand-off each time)
pairs = []
for j, (a, b) in enumerate(zip(big, fo)):
    pairs.append((a, 2 * j - 1 if j else -1))
    pairs.append((b, 2 * j))
run("fc_in->fc_out chain", pairs)
