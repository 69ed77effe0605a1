#!/usr/bin/env python3
"""Small eager launch sequence for ncu: decode GEMVs (C1 shape and the two GPT-J FFN shapes), pure C ABI, no torch."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import load_qmm

qmm = load_qmm()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with qmm.Context(0) as ctx:
    shapes = [(4096, 4096), (16384, 4096), (4096, 16384)]
    ws = []
    for m, k in shapes:
        for i in range(6):          # distinct copies so consecutive launches do not hit L2
            w = qmm.QTensor(ctx, qmm.TYPE_Q4_0, k, m)
            w.set(qmm.random_wire_weights(qmm.TYPE_Q4_0, k, m, seed=i))
            ws.append(w)
    x = ctx.to_device(np.random.default_rng(0).uniform(-1, 1, 16384).astype(np.float32))
    y = ctx.alloc(16384 * 4)
    for _ in range(reps):
        for w in ws:
            ctx.mul_mat_device(w, x.ptr, 1, y.ptr)
    ctx.synchronize()
    print("launches", ctx.launch_count())
