#!/usr/bin/env python3
"""Device-side timeline of a decode chain (b200_ctx_set_trace): where each GEMV launch spends its time and how much
consecutive launches overlap under programmatic dependent launch.  Pure C ABI; graph replay like bench.py."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import load_qmm

qmm = load_qmm()
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mats = [(4096, 4096)] * 4 + [(16384, 4096), (4096, 16384)]
with qmm.Context(0) as ctx:
    ws = []
    for l in range(layers):
        for m, k in mats:
            w = qmm.QTensor(ctx, qmm.TYPE_Q4_0, k, m)
            w.set(qmm.random_wire_weights(qmm.TYPE_Q4_0, k, m, seed=l))
            ws.append(w)
    n_launch = len(ws)
    x = ctx.to_device(np.random.default_rng(0).uniform(-1, 1, 16384).astype(np.float32))
    acts = [ctx.alloc(16384 * 4), ctx.alloc(16384 * 4)]
    trace = ctx.alloc(n_launch * 160 * 8 * 8)
    ctx.lib.b200_memset(ctx.h, trace.ptr, 0, n_launch * 160 * 8 * 8)

    def chain():
        src = x.ptr
        for i, w in enumerate(ws):
            dst = acts[i & 1].ptr
            ctx.mul_mat_device(w, src, 1, dst)
            src = dst
    chain(); ctx.synchronize()
    ctx.set_trace(trace, n_launch)
    ctx.graph_begin(); chain(); g = ctx.graph_end()
    for _ in range(5):
        ctx.graph_launch(g)
    ctx.synchronize()
    t = trace.download(np.uint64, n_launch * 160 * 8).reshape(n_launch, 160, 8).astype(np.int64)
    t0 = t[0, :148, 0].min()
    names = ["entry", "ring primed", "pred done", "act quantized", "first weights", "last row", "stage0 landed", "primed landed"]
    print("launch  shape          " + "  ".join(f"{n:>22s}" for n in names) + "   (ns since first entry: min..max over CTAs)")
    prev_end = None
    for i in range(n_launch):
        m, k = mats[i % len(mats)]
        ctas = min(148, m)
        row = []
        for s in range(8):
            v = t[i, :ctas, s] - t0
            row.append(f"{v.min():9d}..{v.max():9d}")
        end = (t[i, :ctas, 5] - t0).max()
        dur = "" if prev_end is None else f"  +{end - prev_end} ns"
        prev_end = end
        print(f"{i:4d}  {m:6d}x{k:<6d}  " + "  ".join(f"{r:>22s}" for r in row) + dur)
    total = (t[-1, :148, 5].max() - t0)
    nbytes = sum(m * k // 32 * 18 for m, k in mats) * layers
    print(f"total {total} ns for {nbytes/1e6:.1f} MB -> {nbytes/total:.1f} GB/s")
