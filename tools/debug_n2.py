#!/usr/bin/env python3
"""Step-by-step 2-rank check of the row-split + all-gather plumbing (prints after every stage)."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench

def say(*a):
    print(f"[r{os.environ.get('RANK')}] {time.time():.1f}", *a, flush=True)

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
say("init ok")
qmm = bench.load_qmm(); rs = bench.load_rowsplit()
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = qmm.Context(lr, stream=stream.cuda_stream)
pdl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ctx.set_option("pdl", pdl)
mats = [(4096, 4096), (16384, 4096), (4096, 16384)] * 2
ws = []
keep = []
for m, k in mats:
    sp = rs.RowSplit(m, world, rank)
    buf = torch.empty(sp.rows * (k // 32) * 18, dtype=torch.uint8, device=dev); keep.append(buf)
    t = qmm.QTensor(ctx, 2, k, sp.rows, ptr=buf.data_ptr())
    t.set(qmm.random_wire_weights(2, k, m, seed=m + k)[sp.r0:sp.r1])
    ws.append((t, sp, k))
act = [torch.zeros(16384, device=dev), torch.zeros(16384, device=dev)]
x = torch.rand(4096, device=dev) * 2 - 1
dist.broadcast(x, 0)
say("weights ok")

def step():
    src = x
    for i, (t, sp, k) in enumerate(ws):
        dst = act[i & 1]
        rs.gathered_mul_mat(dist, sp, 1, lambda out, ld, t=t, src=src, sp=sp: ctx.mul_mat_device(t, src.data_ptr(), 1, out.data_ptr(), m=sp.rows), dst)
        src = dst
    return src

out = step(); torch.cuda.synchronize(); say("eager step ok", float(out[:4096].abs().sum()))
for _ in range(3): step()
torch.cuda.synchronize(); say("eager x3 ok")
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=stream):
        out = step()
    say("capture ok")
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); say("replay ok", float(out[:4096].abs().sum()))
except Exception as e:
    say("capture failed", type(e).__name__, e)
dist.barrier(); torch.cuda.synchronize(); say("barrier ok")
del g; say("done"); os._exit(0)
