#!/usr/bin/env python3
"""One prefill mul_mat for ncu; pure C ABI.  usage: prof_gemm.py [q4_0|q8_0] [m k n]   (default C2: 11008 4096 512)"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import load_qmm

qmm = load_qmm()
qtype = qmm.TYPE_Q8_0 if (len(sys.argv) > 1 and sys.argv[1] == "q8_0") else qmm.TYPE_Q4_0
m, k, n = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (11008, 4096, 512)
with qmm.Context(0) as ctx:
    w = qmm.QTensor(ctx, qtype, k, m)
    w.set(qmm.random_wire_weights(qtype, k, m, seed=3))
    x = ctx.to_device(np.random.default_rng(0).uniform(-1, 1, (n, k)).astype(np.float32))
    y = ctx.alloc(n * m * 4)
    for _ in range(3):
        ctx.mul_mat_device(w, x.ptr, n, y.ptr)
    ctx.synchronize()
    print("launches", ctx.launch_count())
