#!/usr/bin/env python3
"""Time the C2 prefill mul_mat (m=11008 k=4096 n=512) for Q4_0 and Q8_0, rotating over 6 weight copies (> L2)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import load_qmm
import torch
qmm = load_qmm()
m, k, n = 11008, 4096, 512
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = qmm.Context(0, stream=s.cuda_stream)
for qt, name in ((qmm.TYPE_Q4_0, "q4_0"), (qmm.TYPE_Q8_0, "q8_0")):
    ws = []
    for i in range(6):
        w = qmm.QTensor(ctx, qt, k, m); w.set(qmm.random_wire_weights(qt, k, m, seed=i)); ws.append(w)
    x = torch.rand(n * k, device="cuda") * 2 - 1
    ys = [torch.empty(n * m, device="cuda") for _ in range(6)]
    ctx.reserve_workspace(qt, k, m, n)
    def run():
        for w, y in zip(ws, ys): ctx.mul_mat_device(w, x.data_ptr(), n, y.data_ptr())
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5): run()
    e1.record(s); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    print(name, round(us, 1), "us", round(2.0 * m * n * k / us / 1e6, 1), "TOPS", flush=True)
    for w in ws: w.free()
