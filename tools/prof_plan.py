#!/usr/bin/env python3
"""The GPT-J-6B Q4_0 decode plan (bench.py's workload) launched a few times, for ncu; pure C ABI + torch allocations."""
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
dag = bench.gptj_dag()
host_w, weights, keep = {}, [], []
for name, m, k, _ in dag:
    if (m, k) not in host_w:
        host_w[(m, k)] = qmm.random_wire_weights(2, k, m, seed=1234 + m + k)
    buf = torch.empty(m * (k // 32) * 18, dtype=torch.uint8, device=dev)
    keep.append(buf)
    t = qmm.QTensor(ctx, 2, k, m, ptr=buf.data_ptr())
    t.set(host_w[(m, k)])
    weights.append(t)
x = torch.rand(4096, device=dev) * 2 - 1
lens = [((m + 15) // 16) * 16 for _, m, _, _ in dag]
at = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
out = torch.zeros(int(at[-1]), dtype=torch.float32, device=dev)
args = [ctx.make_args(weights[i], x.data_ptr() if s < 0 else out.data_ptr() + int(at[s]) * 4, 1, out.data_ptr() + int(at[i]) * 4)
        for i, (_, m, k, s) in enumerate(dag)]
plan = ctx.plan_create(args)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    ctx.plan_launch(plan)
ctx.synchronize()
print("launches", ctx.launch_count(), "logits finite", bool(torch.isfinite(out[int(at[-2]):int(at[-2]) + 50400]).all()))
