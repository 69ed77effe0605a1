#!/usr/bin/env python3
"""A/B of the prefill GEMM paths on BASELINE.json's C2 (m = 11008, k = 4096, n = 512; Q4_0 and Q8_0), one process:
  B200_GEMM_F16 unset  the exact kernel (tcgen05 kind::i8 per 32-wide k-block, fp32 scaling on CUDA cores)   [shipped]
  B200_GEMM_F16=1      fp16 operands materialised in scratch, tcgen05 kind::f16, fp32 accumulation            [experimental]
  B200_GEMM_F16=2      weights dequantized inside the kernel into the swizzled operand tile                   [experimental]
Per mode: NMSE against the exact kernel, us per mul_mat (CUDA events over 6 rotating weight copies > L2, quantize_q8_0 of the
activations included), int8-equivalent TOPS.  Usage on the GPU box: python tools/ab_gemm.py [m k n]"""
import os
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
m, k, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (11008, 4096, 512)
WIRE = {2: 18, 8: 34}


def nmse(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300))


for qtype, name in ((2, "q4_0"), (8, "q8_0")):
    nrot = 6
    wire = qmm.random_wire_weights(qtype, k, m, seed=9)
    bufs, ts, ys = [], [], []
    for i in range(nrot):
        b = torch.empty(m * (k // 32) * WIRE[qtype], dtype=torch.uint8, device=dev)
        t = qmm.QTensor(ctx, qtype, k, m, ptr=b.data_ptr())
        t.set(wire)
        bufs.append(b); ts.append(t)
        ys.append(torch.empty(n * m, dtype=torch.float32, device=dev))
    x = torch.rand(n * k, dtype=torch.float32, device=dev) * 2 - 1
    exact = None
    for mode in (None, "1", "2"):
        if mode is None:
            os.environ.pop("B200_GEMM_F16", None)
        else:
            os.environ["B200_GEMM_F16"] = mode
        ctx.reserve_workspace(qtype, k, m, n)
        for y in ys:
            y.zero_()

        def run():
            for t, y in zip(ts, ys):
                ctx.mul_mat_device(t, x.data_ptr(), n, y.data_ptr(), flags=qmm.MM_FORCE_GEMM)

        run()
        torch.cuda.synchronize()
        got = ys[0].cpu().numpy()
        if mode is None:
            exact = got
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (reps * nrot) * 1e3
        tops = 2.0 * m * n * k / (us * 1e-6) / 1e12
        print(f"{name} m={m} k={k} n={n}  B200_GEMM_F16={mode or '-':1s}  {us:8.1f} us/mul_mat  {tops:7.1f} TOPS-equivalent  "
              f"finite {bool(np.isfinite(got).all())}  nmse vs exact {nmse(got, exact):.3e}", flush=True)
    del bufs, ts, ys
