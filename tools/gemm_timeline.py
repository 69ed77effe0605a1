#!/usr/bin/env python3
"""Device-side timeline of one gemm_f16_pair_kernel launch (b200_ctx_set_trace): per CTA eight %globaltimer stamps -- 0 start, 1 first MMA
issued, 2 MMAs of the first unit issued, 3 first accumulator complete, 4 first unit stored, 5 last unit (k-slice) stored, 6 all k-slices
arrived, 7 end.  Prints min / median / max over the CTAs in us since the earliest start.  usage: gemm_timeline.py [q4_0|q8_0] [m k n]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import load_qmm

qmm = load_qmm()
qtype = qmm.TYPE_Q8_0 if (len(sys.argv) > 1 and sys.argv[1] == "q8_0") else qmm.TYPE_Q4_0
m, k, n = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (11008, 4096, 512)
ablate = int(sys.argv[5]) if len(sys.argv) > 5 else 0
NAMES = ["start", "first MMA issued", "unit 0: MMAs issued", "unit 0: accumulator complete", "unit 0: stored", "last unit: stored", "all k-slices arrived", "end"]
with qmm.Context(0) as ctx:
    if ablate:
        raise SystemExit("the ablation switches belong to tools/experiments/b200_gemm_f16_tmem_a.cu (not in the shipped kernel)")
    w = qmm.QTensor(ctx, qtype, k, m)
    w.set(qmm.random_wire_weights(qtype, k, m, seed=3))
    x = ctx.to_device(np.random.default_rng(0).uniform(-1, 1, (n, k)).astype(np.float32))
    y = ctx.alloc(n * m * 4)
    for _ in range(3):
        ctx.mul_mat_device(w, x.ptr, n, y.ptr)
    ctx.synchronize()
    nl = 4 if not (ablate & 4) else 2
    tb = ctx.alloc(nl * 160 * 8 * 8)
    ctx.memset(tb, 0, nl * 160 * 8 * 8) if hasattr(ctx, "memset") else tb.upload(np.zeros(nl * 160 * 8, np.uint64))
    ctx.set_trace(tb, nl)
    for _ in range(nl if not (ablate & 4) else 1):
        ctx.mul_mat_device(w, x.ptr, n, y.ptr)
    ctx.synchronize()
    ctx.set_trace(None, 0)
    t = tb.download(np.uint64, nl * 160 * 8).reshape(nl, 160, 8).astype(np.int64)
    # the GEMV-class kernels (quantize) do not take slots; every slot is one GEMM launch
    if ablate & 4:
        t2 = t[1]
        mma = t2[0::2][t2[0::2, 3] > 0]
        deq = t2[1::2][t2[1::2, 5] > 0]
        if mma.size:
            c = mma[:, 3].astype(np.float64)
            print(f"MMA thread, cycles per k-step (median over {len(mma)} leaders, {int(np.median(c))} k-steps): wait X' {np.median(mma[:, 0] / c):7.1f}  wait W' {np.median(mma[:, 1] / c):7.1f}  issue+commit {np.median(mma[:, 2] / c):7.1f}   SM clock {np.median(mma[:, 4] / np.maximum(mma[:, 5], 1)):5.3f} GHz, {np.median(mma[:, 4] / c):6.1f} cycles per k-step in all")
        if deq.size:
            c = deq[:, 5].astype(np.float64)
            print(f"dequant warp (group 0), cycles per own k-step (median over {len(deq)} CTAs, {int(np.median(c))} k-steps): wait raw {np.median(deq[:, 0] / c):7.1f}  lds+convert {np.median(deq[:, 1] / c):7.1f}  "
                  f"flush prev st {np.median(deq[:, 2] / c):7.1f}  wait W' free {np.median(deq[:, 3] / c):7.1f}  st issue {np.median(deq[:, 4] / c):7.1f}")
    for l in range(0 if ablate & 4 else nl - 1, 1 if ablate & 4 else nl):
        tl = t[l]
        live = tl[:, 0] > 0
        t0 = tl[live, 0].min()
        print(f"launch {l}: {int(live.sum())} CTAs, {qmm.TYPE_NAMES[qtype]} m={m} k={k} n={n} ablate={ablate}")
        for s in range(8):
            v = tl[live, s]
            v = v[v > 0] - t0
            if v.size:
                print(f"  {s} {NAMES[s]:32s} n={v.size:4d}  min {v.min() / 1e3:8.2f}  median {np.median(v) / 1e3:8.2f}  max {v.max() / 1e3:8.2f} us")
