#!/usr/bin/env python3
"""A/B of the prefill GEMM paths, one process:
  gemm_exact = 1   tcgen05 kind::i8 per 32-wide k-block, exact int32 partials, fp32 scaling on CUDA cores  (b200_gemm_tc.cu)
  gemm_exact = 0   the default: fp16 operands (weights dequantized in the kernel), tcgen05 kind::f16, cta_group::2 (b200_gemm_f16.cu)
Per shape and type: NMSE against the exact kernel, us per mul_mat (CUDA events over rotating weight copies > L2, activation
quantization included), TFLOP/s.  Usage on the GPU box: python tools/ab_gemm.py [m k n]"""
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
shapes = [tuple(int(a) for a in sys.argv[1:4])] if len(sys.argv) >= 4 else [(11008, 4096, 512), (4096, 4096, 512), (16384, 4096, 512), (4096, 16384, 512), (50400, 4096, 512),
                                                                         (28672, 8192, 512), (11008, 4096, 128), (11008, 4096, 2048)]
WIRE = {2: 18, 8: 34}


def nmse(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300))


for (m, k, n) in shapes:
    for qtype, name in ((2, "q4_0"), (8, "q8_0")):
        per = m * (k // 32) * WIRE[qtype] + n * m * 4
        nrot = max(2, min(8, int(300e6 // per) + 1))
        wire = qmm.random_wire_weights(qtype, k, m, seed=9)
        bufs, ts, ys = [], [], []
        for i in range(nrot):
            b = torch.empty(m * (k // 32) * WIRE[qtype], dtype=torch.uint8, device=dev)
            t = qmm.QTensor(ctx, qtype, k, m, ptr=b.data_ptr())
            t.set(wire)
            bufs.append(b); ts.append(t)
            ys.append(torch.empty(n * m, dtype=torch.float32, device=dev))
        x = torch.rand(n * k, dtype=torch.float32, device=dev) * 2 - 1
        ctx.reserve_workspace(qtype, k, m, n)
        exact = None
        for mode in (1, 0):
            if mode == 1 and (m * n * k > 3e11 or len(shapes) > 1 and (m, k, n) != shapes[0]):
                continue        # the exact kernel is only timed on the first shape
            ctx.set_option("gemm_exact", mode)
            for y in ys:
                y.zero_()

            def run():
                for t, y in zip(ts, ys):
                    ctx.mul_mat_device(t, x.data_ptr(), n, y.data_ptr())

            run()
            ctx.synchronize()
            got = ys[0].cpu().numpy()
            if mode == 1:
                exact = got
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record(stream)
            for _ in range(reps):
                run()
            e1.record(stream)
            ctx.synchronize()
            us = e0.elapsed_time(e1) / (reps * nrot) * 1e3
            tf = 2.0 * m * n * k / (us * 1e-6) / 1e12
            err = f"nmse vs exact {nmse(got, exact):.3e}" if exact is not None else ""
            print(f"{name} m={m} k={k} n={n}  {'exact int8' if mode else 'fp16 pair '}  {us:8.1f} us/mul_mat  {tf:7.1f} TFLOP/s  finite {bool(np.isfinite(got).all())}  {err}", flush=True)
        ctx.set_option("gemm_exact", 0)
        del bufs, ts, ys
