#!/usr/bin/env python3
"""mul_mat latency against the number of activation columns n, GEMV (column chunks of <= 8) against the tensor-core GEMM, per shape and type:
where the default switch between the two (context option gemv_max_n) should sit.  Rotates over weight copies larger than L2."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import load_qmm
import torch
qmm = load_qmm()
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = qmm.Context(0, stream=s.cuda_stream)
shapes = [(4096, 4096), (16384, 4096), (4096, 16384)]
ns = [1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 128]
for qt, name in ((qmm.TYPE_Q4_0, "q4_0"), (qmm.TYPE_Q8_0, "q8_0")):
    for m, k in shapes:
        copies = max(2, int(400e6 // (m * k // 32 * (18 if qt == qmm.TYPE_Q4_0 else 34))) + 1)
        ws = []
        for i in range(copies):
            w = qmm.QTensor(ctx, qt, k, m); w.set(qmm.random_wire_weights(qt, k, m, seed=i)); ws.append(w)
        x = torch.rand(128 * k, device="cuda") * 2 - 1
        y = torch.empty(128 * m, device="cuda")
        ctx.reserve_workspace(qt, k, m, 128)
        row = []
        for n in ns:
            res = {}
            for label, flags in (("gemv", qmm.MM_FORCE_GEMV), ("gemm", qmm.MM_FORCE_GEMM), ("default", 0)):
                def run():
                    for w in ws: ctx.mul_mat_device(w, x.data_ptr(), n, y.data_ptr(), flags=flags)
                try:
                    run(); torch.cuda.synchronize()
                except Exception as e:
                    res[label] = float("nan"); continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record(s)
                for _ in range(reps): run()
                e1.record(s); torch.cuda.synchronize()
                res[label] = e0.elapsed_time(e1) / (reps * len(ws)) * 1e3
            row.append((n, res))
        print(f"{name} m={m} k={k}: " + "  ".join(f"n={n}: gemv {r['gemv']:.1f} gemm {r['gemm']:.1f} dflt {r['default']:.1f}" for n, r in row), flush=True)
        for w in ws: w.free()
