#!/usr/bin/env python3
"""Repeat the fp16 prefill GEMM on one shape and look for run-to-run differences (races): every run is compared with the exact int8
kernel's result; for a bad run the wrong elements are located (tile, k-slice owner columns).  usage: stress_gemm.py [q4_0|q8_0] m k n [runs]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

qmm = bench.load_qmm()
qtype = 8 if sys.argv[1] == "q8_0" else 2
m, k, n = (int(a) for a in sys.argv[2:5])
runs = int(sys.argv[5]) if len(sys.argv) > 5 else 40
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = qmm.Context(0, stream=stream.cuda_stream)
wire = qmm.random_wire_weights(qtype, k, m, seed=m + k)
t = qmm.QTensor(ctx, qtype, k, m)
t.set(wire)
x = torch.rand(n * k, dtype=torch.float32, device=dev) * 2 - 1
y = torch.empty(n * m, dtype=torch.float32, device=dev)
ctx.set_option("gemm_exact", 1)
ctx.mul_mat_device(t, x.data_ptr(), n, y.data_ptr(), flags=qmm.MM_FORCE_GEMM)
ctx.synchronize()
exact = y.clone().view(n, m)
ctx.set_option("gemm_exact", 0)
scale = float(exact.pow(2).mean().sqrt())
first = None
nbad = 0
for r in range(runs):
    y.fill_(float("nan"))
    ctx.mul_mat_device(t, x.data_ptr(), n, y.data_ptr())
    if r % 3 == 0:
        ctx.mul_mat_device(t, x.data_ptr(), n, y.data_ptr())     # back to back
    ctx.synchronize()
    got = y.view(n, m)
    if first is None:
        first = got.clone()
    err = (got - exact).abs()
    wrong = (err > 0.02 * scale) | ~torch.isfinite(got)
    same = bool(torch.equal(got, first))
    if wrong.any() or not same:
        nbad += 1
        idx = wrong.nonzero()
        cols, rows = idx[:, 0], idx[:, 1]
        print(f"run {r}: {int(wrong.sum())} wrong elements, equal to first run: {same}; cols {int(cols.min()) if len(cols) else -1}..{int(cols.max()) if len(cols) else -1} "
              f"rows {int(rows.min()) if len(rows) else -1}..{int(rows.max()) if len(rows) else -1}; distinct row tiles(256) {sorted(set((rows // 256).tolist()))[:12]} "
              f"distinct row%256//32 {sorted(set(((rows % 256) // 32).tolist()))} distinct cols//32 {sorted(set((cols // 32).tolist()))[:16]} max err {float(err[torch.isfinite(err)].max()) / scale:.3f} x rms")
print(f"{sys.argv[1]} m={m} k={k} n={n}: {nbad} bad runs of {runs}")
