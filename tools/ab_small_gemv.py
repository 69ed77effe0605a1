#!/usr/bin/env python3
"""Decode GEMV on SMALL matrices (GPT-2 117M shapes, L2-resident weights): the streaming kernel (persistent CTAs + bulk-copy ring, built for
9-150 MB matrices) against the generic warp-per-row-pair kernel, us per launch over a chain of launches on one stream, plus the two glue
kernels that dominate a GPT-2 decode step.  Usage on the GPU box: python tools/ab_small_gemv.py"""
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
WIRE = {2: 18, 8: 34}


def timed(fn, reps=200):
    fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (m, k) in [(2304, 768), (768, 768), (3072, 768), (768, 3072), (50257, 768), (4096, 4096), (1024, 1024), (8192, 2048), (16384, 4096)]:
    for qtype, name in ((2, "q4_0"), (8, "q8_0")):
        nrot = 4
        wire = qmm.random_wire_weights(qtype, k, m, seed=9)
        ts, ys, keep = [], [], []
        for i in range(nrot):
            b = torch.empty(m * (k // 32) * WIRE[qtype], dtype=torch.uint8, device=dev)
            t = qmm.QTensor(ctx, qtype, k, m, ptr=b.data_ptr())
            t.set(wire)
            keep.append(b); ts.append(t)
            ys.append(torch.empty(m, dtype=torch.float32, device=dev))
        x = torch.rand(k, dtype=torch.float32, device=dev) * 2 - 1
        res = {}
        out = {}
        for mode in (1, 0):
            ctx.set_option("gemv_stream", mode)

            def run():
                for t, y in zip(ts, ys):
                    ctx.mul_mat_device(t, x.data_ptr(), 1, y.data_ptr())

            res[mode] = timed(run, 100) / nrot
            out[mode] = ys[0].cpu().numpy().copy()
        ctx.set_option("gemv_stream", 1)
        mb = m * (k // 32) * WIRE[qtype] / 1e6
        print(f"{name} m={m:6d} k={k:5d} ({mb:7.2f} MB)  stream {res[1]:7.2f} us   generic {res[0]:7.2f} us   same bits {bool(np.array_equal(out[0], out[1]))}", flush=True)
