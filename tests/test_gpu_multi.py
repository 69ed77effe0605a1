"""Row split across TWO GPUs of one box (skipped when fewer are visible; the driver's GPU test box may have one): one process per
GPU, torch.distributed for the plumbing.  Everything is compared with the ORACLE on the unsplit matrices and, bit for bit, with
the single-GPU result (a row of dst is computed by the same code whichever rank owns it).
  * the row-split decode plan (b200_plan_create with a b200_plan_split: tagged NVLink peer stores, published vectors included);
  * the per-launch fused gather (b200_mul_mat_gather);
  * prefill: every rank's dst slice [n][rows] through the tensor-core GEMM, NCCL all-gather, strided scatter
    (rowsplit.gathered_mul_mat; the issue the reference notes at src/ggml-cuda.cu:1592-1608)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
Q4_0, Q8_0 = 2, 8


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")]

# a GPT-J-like block pair at reduced width + a head that does not divide evenly: [(m, k, src)]
NODES = [(512, 512, -1), (16384, 512, -1), (512, 512, -1), (512, 512, -1), (512, 512, 0), (512, 16384, 1),
         (512, 512, 5), (1000, 512, 5), (257, 512, 6)]


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from conftest import Oracle, load_qmm, nmse, MUL_MAT_NMSE_TOL, F16_GEMM_NMSE
    from test_rowsplit_gloo import load_rowsplit
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    qmm, rs, oracle = load_qmm(), load_rowsplit(), Oracle()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = qmm.Context(rank, stream=stream.cuda_stream)
    ok, notes = True, []
    rng = np.random.default_rng(77)                                   # same data on every rank
    wires = [oracle.quantize_weights(Q4_0, rng.uniform(-1, 1, (m, k)).astype(np.float32) * np.float32(np.sqrt(3.0 / k))) for m, k, _ in NODES]
    x = rng.uniform(-1, 1, NODES[0][1]).astype(np.float32)
    # ---- the oracle on the unsplit graph, node by node on the device's own inputs (per-node parity), and chained
    splits = [rs.RowSplit(m, world, rank) for m, _, _ in NODES]
    ws = []
    for (m, k, _), sp, wire in zip(NODES, splits, wires):
        t = qmm.QTensor(ctx, Q4_0, k, max(sp.rows, 1))
        if sp.rows > 0:
            t.set(wire[sp.r0:sp.r1])
        ws.append(t)
    xd = ctx.to_device(x)
    lens = [((m + 31) // 32) * 32 for m, _, _ in NODES]
    at = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = ctx.alloc(int(at[-1]) * 4)
    ctx._check(ctx.lib.b200_memset(ctx.h, out.ptr, 0xff, out.nbytes))
    args = []
    for i, ((m, k, s), sp) in enumerate(zip(NODES, splits)):
        a = ctx.make_args(ws[i], xd.ptr if s < 0 else out.ptr + int(at[s]) * 4, 1, out.ptr + int(at[i]) * 4, m=sp.rows)
        a.ne02 = a.ne03 = 1
        a.flags |= qmm.MM_EXPORT                                      # every node complete on every rank: all of them are checked
        args.append(a)
    ps = rs.plan_split(qmm.PlanSplit, splits, world, rank)
    arena = ctx.alloc(ctx.plan_arena_bytes(args, ps))
    ctx._check(ctx.lib.b200_memset(ctx.h, arena.ptr, 0, arena.nbytes))
    handles = [None] * world
    dist.all_gather_object(handles, ctx.ipc_export(arena.ptr))
    for r in range(world):
        ps.peer_arena[r] = arena.ptr if r == rank else ctx.ipc_import(handles[r])
    dist.barrier()
    for min_k in (256, 0):                                            # every possible vector published / none
        ctx.set_option("plan_pub_min_k", min_k)
        plan = ctx.plan_create(args, ps)
        for _ in range(3):
            ctx.plan_launch(plan)
        ctx.synchronize()
        got = [out.download(np.float32, NODES[i][0], offset=int(at[i]) * 4) for i in range(len(NODES))]
        for i, (m, k, s) in enumerate(NODES):
            src = x if s < 0 else got[s]
            ref = oracle.mul_mat(Q4_0, wires[i], k, m, 1, 1, src.reshape(1, 1, 1, k))[0, 0, 0]
            e = nmse(got[i], ref)
            if not (np.isfinite(got[i]).all() and e <= MUL_MAT_NMSE_TOL and e <= 1e-9):
                ok = False
                notes.append(f"plan(min_k={min_k}) node {i}: nmse {e}")
        # every rank holds the same bits
        mine = torch.from_numpy(np.concatenate(got)).to(dev)
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        if not all(torch.equal(both[0], b) for b in both):
            ok = False
            notes.append(f"plan(min_k={min_k}): ranks disagree")
        ctx.plan_destroy(plan)
        dist.barrier()
    ctx.set_option("plan_pub_min_k", 4096)
    # ---- single-GPU bits: rank 0 runs the unsplit graph node by node and compares
    if rank == 0:
        full = [qmm.QTensor(ctx, Q4_0, k, m) for m, k, _ in NODES]
        for t, wire in zip(full, wires):
            t.set(wire)
        o1 = ctx.alloc(int(at[-1]) * 4)
        for i, (m, k, s) in enumerate(NODES):
            ctx.mul_mat_device(full[i], xd.ptr if s < 0 else o1.ptr + int(at[s]) * 4, 1, o1.ptr + int(at[i]) * 4)
        ctx.synchronize()
        for i, (m, _, _) in enumerate(NODES):
            if not np.array_equal(got[i], o1.download(np.float32, m, offset=int(at[i]) * 4)):
                ok = False
                notes.append(f"node {i}: row-split plan differs from the single-GPU bits")
    # ---- prefill: n = 96 columns, m = 1000 (uneven) and m = 512, through the GEMM + NCCL all-gather + strided scatter
    for (m, k, n, qtype) in ((1000, 512, 96, Q4_0), (512, 1024, 40, Q8_0)):
        wire = oracle.quantize_weights(qtype, rng.uniform(-1, 1, (m, k)).astype(np.float32))
        xs = rng.uniform(-1, 1, (n, k)).astype(np.float32)
        sp = rs.RowSplit(m, world, rank)
        t = qmm.QTensor(ctx, qtype, k, max(sp.rows, 1))
        if sp.rows > 0:
            t.set(wire[sp.r0:sp.r1])
        xt = torch.from_numpy(xs).to(dev)
        dst = torch.zeros(n * m, dtype=torch.float32, device=dev)
        staging = torch.zeros(n * sp.chunk, dtype=torch.float32, device=dev)
        gathered = torch.zeros(world * n * sp.chunk, dtype=torch.float32, device=dev)

        def compute_slice(o, ld, t=t, sp=sp, xt=xt, n=n):
            # the slice is a dense [n][rows] matrix (ld == chunk may exceed rows: compute into a dense temp, then place)
            tmp = torch.empty(n * sp.rows, dtype=torch.float32, device=dev)
            ctx.mul_mat_device(t, xt.data_ptr(), n, tmp.data_ptr(), m=sp.rows)
            o.view(n, ld)[:, :sp.rows].copy_(tmp.view(n, sp.rows))
        res = rs.gathered_mul_mat(dist, sp, n, compute_slice, dst, staging, gathered)
        torch.cuda.synchronize()
        ref = oracle.mul_mat(qtype, wire, k, m, 1, 1, xs[None, None])[0, 0]
        e = nmse(res.cpu().numpy(), ref)
        if not (e <= F16_GEMM_NMSE):
            ok = False
            notes.append(f"prefill row split m={m} n={n}: nmse {e}")
    ret[rank] = (ok, notes)
    dist.barrier()
    dist.destroy_process_group()


def test_row_split_on_two_gpus():
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert all(ret.get(r, (False, ["no result"]))[0] for r in range(2)), dict(ret)


def test_split_buffer_type_through_the_reference_api():
    """ggml_backend_cuda_split_buffer_type (src/ggml-cuda.h:28-29) dropped in: the weights of a ggml graph live in the split buffer
    type (rows divided across both GPUs of this process), the graph is computed through ggml_backend_graph_compute on the B200
    backend, and every node matches the reference CPU backend.  Decode graphs go down as ONE row-split persistent launch per
    device; n = 3 takes the slice-by-slice path (dst rows copied into place)."""
    import ctypes as C
    from conftest import Oracle, nmse, MUL_MAT_NMSE_TOL, WIRE
    from test_gpu_dropin_graph import load, make_graph, node_outputs, NODES, SHAPES, E, vp
    oracle = Oracle()
    cpu = load("libref_shim.so")
    gpu = load("libdropin_shim.so")
    gpu.ref_dag_create_split.restype = vp
    gpu.ref_select_backend(b"B2000")
    for qtype in (Q4_0, Q8_0):
        for ncols in (1, 3):
            rng = np.random.default_rng(31 + qtype + ncols)
            hc = make_graph(cpu, qtype, NODES, SHAPES, ncols)
            n = len(NODES)
            hg = gpu.ref_dag_create_split(qtype, n, (C.c_int * n)(*[w for w, _ in NODES]), (C.c_int * n)(*[s for _, s in NODES]), len(SHAPES),
                                          (C.c_int64 * len(SHAPES))(*[k for k, _ in SHAPES]), (C.c_int64 * len(SHAPES))(*[m for _, m in SHAPES]),
                                          C.c_int64(ncols), 4)
            assert hg, "ref_dag_create_split failed"
            hg = vp(hg)
            try:
                for j, (k, m) in enumerate(SHAPES):
                    w = oracle.quantize_weights(qtype, rng.uniform(-1, 1, (m, k)).astype(np.float32) * (2.0 / np.sqrt(k)))
                    cpu.ref_chain_set_weight(hc, j, w.ctypes.data_as(vp))
                    gpu.ref_chain_set_weight(hg, j, w.ctypes.data_as(vp))
                for rep in range(3):
                    x = rng.uniform(-1, 1, (ncols, E)).astype(np.float32)
                    cpu.ref_chain_set_x(hc, x.ctypes.data_as(vp))
                    gpu.ref_chain_set_x(hg, x.ctypes.data_as(vp))
                    cpu.ref_chain_compute(hc)
                    gpu.ref_chain_compute(hg)
                    want, got = node_outputs(cpu, hc), node_outputs(gpu, hg)
                    for i, (a, b) in enumerate(zip(got, want)):
                        assert np.isfinite(a).all(), f"node {i}"
                        assert nmse(a, b) <= MUL_MAT_NMSE_TOL, f"type {qtype} ncols {ncols} rep {rep} node {i}: nmse {nmse(a, b)}"
                gpu.ref_chain_plan_launches.restype = C.c_int64
                gpu.ref_chain_plan_launches.argtypes = [vp]
                plans = gpu.ref_chain_plan_launches(hg)
                assert plans == (3 if ncols == 1 else 0), f"decode graph on split weights: {plans} plan launches"
            finally:
                cpu.ref_chain_free(hc)
                gpu.ref_chain_free(hg)
