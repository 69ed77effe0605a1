#!/usr/bin/env python3
"""BASELINE.json configs[3], second half: the 512-token prefill of the GPT-J-6B mul_mat graph ROW-SPLIT across N GPUs.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/prefill_rowsplit.py
Every rank computes its row slice of every node with the tensor-core GEMM ([n][rows] dense), the slices are re-assembled on
every rank by ONE NCCL all-gather per node + a strided scatter (rowsplit.gathered_mul_mat; the layout issue the reference notes
at src/ggml-cuda.cu:1592-1608).  Checked against the oracle port on sampled rows of the logits (device inputs of the last
node); timed with CUDA events, max over ranks.  One JSON line from rank 0."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
qmm, rs = bench.load_qmm(), bench.load_rowsplit()
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = qmm.Context(local, stream=stream.cuda_stream)
wl = bench.Workload("gptj")
dag, npf = wl.nodes, 512
host_w = {(m, k): qmm.random_wire_weights(2, k, m, seed=wl.seed(m, k)) for (m, k) in wl.shapes}
weights, keep = [], []
for name, m, k, _ in dag:
    sp = rs.RowSplit(m, world, rank)
    buf = torch.empty(max(sp.rows, 1) * (k // 32) * 18, dtype=torch.uint8, device=dev)
    keep.append(buf)
    t = qmm.QTensor(ctx, 2, k, max(sp.rows, 1), ptr=buf.data_ptr())
    if sp.rows > 0:
        t.set(host_w[(m, k)][sp.r0:sp.r1])
    weights.append((t, sp, k))
    ctx.reserve_workspace(2, k, max(sp.rows, 1), npf)
xin = torch.rand(npf * wl.x_len, dtype=torch.float32, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 2 - 1
if world > 1:
    dist.broadcast(xin, 0)
blk = [[torch.empty(npf * m, dtype=torch.float32, device=dev) for _, m, _, _ in wl.block] for _ in range(2)]
head = torch.empty(npf * wl.out_len, dtype=torch.float32, device=dev)
cmax = max(sp.chunk for _, sp, _ in weights)
staging = torch.empty(npf * cmax, dtype=torch.float32, device=dev)
gathered = torch.empty(world * npf * cmax, dtype=torch.float32, device=dev)
n_nodes = len(dag)


def out_of(i):
    return head if i == n_nodes - 1 else blk[(i // wl.per_block) & 1][i % wl.per_block]


def prefill():
    for i, (t, sp, k) in enumerate(weights):
        src = dag[i][3]
        sptr = xin.data_ptr() if src < 0 else out_of(src).data_ptr()
        if world == 1:
            ctx.mul_mat_device(t, sptr, npf, out_of(i).data_ptr())
        else:
            # (ld == chunk == rows except on a short last rank: this graph's m all divide by 2, 4 and 8 except the head at 8 -> 6300)
            rs.gathered_mul_mat(dist, sp, npf, lambda o, ld, t=t, sptr=sptr, sp=sp: ctx.mul_mat_device(t, sptr, npf, o.data_ptr(), m=sp.rows) if ld == sp.rows else
                                o.view(npf, ld)[:, :sp.rows].copy_(_dense(t, sptr, sp)), out_of(i), staging, gathered)


def _dense(t, sptr, sp):
    tmp = torch.empty(npf * sp.rows, dtype=torch.float32, device=dev)
    ctx.mul_mat_device(t, sptr, npf, tmp.data_ptr(), m=sp.rows)
    return tmp.view(npf, sp.rows)


def timed(reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        prefill()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms


prefill()
torch.cuda.synchronize()
ms = float(np.median([timed(3) for _ in range(5)]))
cols = [0, 1, 255, 511]
hin = out_of(dag[-1][3]).view(npf, wl.x_len)[cols].cpu().numpy()
rows = np.unique(np.random.default_rng(3).integers(0, wl.out_len, 48))
got = head.view(npf, wl.out_len)[cols][:, torch.from_numpy(rows).to(dev)].cpu().numpy()
if rank == 0:
    ref = bench.oracle_rows(host_w[(wl.out_len, wl.x_len)][rows], wl.x_len, hin)
    err = bench.nmse(got, ref)
    ops = 2.0 * npf * sum(m * k for _, m, k, _ in dag)
    print(json.dumps({"what": "GPT-J-6B Q4_0 512-token prefill mul_mat graph, row-split + NCCL all-gather + strided scatter", "n_gpus": world,
                      "ms": round(ms, 3), "prompt_tokens/s": round(npf * 1000.0 / ms, 1), "TFLOP/s": round(ops / (ms * 1e-3) / 1e12, 1),
                      "lm_head_sample_vs_oracle_nmse": float(f"{err:.3e}"), "ok": bool(err <= bench.NMSE_TOL)}))
if world > 1:
    dist.barrier()
    os._exit(0)
