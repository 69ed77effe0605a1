"""Row-split + all-gather host logic (ggml-imax_b200/rowsplit.py) on CPU: world_size 2 (and 3) over gloo, the per-rank
slice computed by the ORACLE, the assembled result compared with the unsplit oracle result bit for bit.  Covers the
n == 1 in-place path, the n > 1 staging/permute path, uneven m (padding) and a rank that owns no rows."""
import importlib.util
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
Q4_0, Q8_0 = 2, 8


def load_rowsplit():
    spec = importlib.util.spec_from_file_location("ggml_imax_b200_rowsplit", ROOT / "ggml-imax_b200" / "rowsplit.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ggml_imax_b200_rowsplit"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_row_ranges():
    rs = load_rowsplit()
    for m, world in [(4096, 8), (50400, 8), (50257, 2), (5, 8), (1, 2), (28672, 4)]:
        covered = []
        for r in range(world):
            s = rs.RowSplit(m, world, r)
            assert s.ranges()[r] == (s.r0, s.r1) and 0 <= s.rows <= s.chunk
            covered.extend(range(s.r0, s.r1))
        assert covered == list(range(m))
        assert rs.RowSplit(m, world, 0).padded_m >= m
    assert rs.message_bytes(4096, 1, 8) == 2048 and rs.message_bytes(50400, 1, 8) == 25200


def _worker(rank, world, port, cases, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import Oracle
    rs = load_rowsplit()
    oracle = Oracle()
    ok = True
    for qtype, m, k, n in cases:
        rng = np.random.default_rng(1000 + m + k + n)        # same data on every rank
        wire = oracle.quantize_weights(qtype, rng.uniform(-1, 1, (m, k)))
        x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
        ref = oracle.mul_mat(qtype, wire, k, m, 1, 1, x[None, None])[0, 0]
        split = rs.RowSplit(m, world, rank)

        def compute_slice(out, ld, split=split, wire=wire, x=x, qtype=qtype, k=k, n=n):
            y = oracle.mul_mat(qtype, np.ascontiguousarray(wire[split.r0:split.r1]), k, split.rows, 1, 1, x[None, None])[0, 0]
            o = out.view(-1)
            for j in range(n):
                o[j * ld: j * ld + split.rows] = torch.from_numpy(y[j])

        dst = torch.full((max(n * split.padded_m, n * m),), float("nan"))
        staging = torch.zeros(n * split.chunk)
        gathered = torch.zeros(world * n * split.chunk)
        out = rs.gathered_mul_mat(dist, split, n, compute_slice, dst, staging, gathered)
        ok = ok and out.shape == (n, m) and np.array_equal(out.numpy().view(np.uint32), ref.view(np.uint32))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gathered_mul_mat_gloo(world):
    cases = [(Q4_0, 64, 128, 1), (Q4_0, 65, 128, 1), (Q8_0, 37, 64, 4), (Q4_0, 2, 32, 1), (Q8_0, 128, 256, 1), (Q4_0, 50, 96, 9)]
    port = 29500 + (os.getpid() % 400) + world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, cases, ret), nprocs=world, join=True)
    assert all(ret.get(r, False) for r in range(world)), dict(ret)


# ---- row-split decode plan: host logic of every rank (rowsplit.plan_split + b200_plan_analyze), the exchange emulated --------

def load_qmm():
    spec = importlib.util.spec_from_file_location("ggml_imax_b200_qmm", ROOT / "ggml-imax_b200" / "qmm.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ggml_imax_b200_qmm"] = mod
    spec.loader.exec_module(mod)
    return mod


def _plan_worker(rank, world, port, nodes, ret):
    """nodes: [(m, k, src)].  Every rank describes ITS slices to b200_plan_analyze (no device needed), then the plan's
    semantics are played with the oracle: per op every rank computes its rows from the full src1 vector and the slices
    are exchanged (what the tagged NVLink stores of the kernel do); the last vector must equal the unsplit chain bit for bit."""
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import Oracle
    rs, qmm, oracle = load_rowsplit(), load_qmm(), Oracle()
    lib = qmm.load_library()
    rng = np.random.default_rng(77)                                   # same data on every rank
    wires = [oracle.quantize_weights(Q4_0, rng.uniform(-1, 1, (m, k)).astype(np.float32) * np.float32(np.sqrt(3.0 / k))) for m, k, _ in nodes]
    x = rng.uniform(-1, 1, nodes[0][1]).astype(np.float32)
    splits = [rs.RowSplit(m, world, rank) for m, _, _ in nodes]
    ps = rs.plan_split(qmm.PlanSplit, splits, world, rank)
    args = (qmm.MulMatArgs * len(nodes))()
    dst = [0x10000000 + i * 0x400000 for i in range(len(nodes))]
    for i, (m, k, src) in enumerate(nodes):
        a = args[i]
        a.type = Q4_0
        a.src0_dev = 0x40000000 + i * 0x4000000
        a.src0_nblocks_total = max(splits[i].rows, 1) * (k // 32)
        a.ne00, a.ne01, a.ne02, a.ne03 = k, splits[i].rows, 1, 1
        a.src1_dev = 0x0F000000 if src < 0 else dst[src]
        a.ne11 = a.ne12 = a.ne13 = 1
        a.nb11 = a.nb12 = a.nb13 = k * 4
        a.dst_dev = dst[i]
        if i == len(nodes) - 1:
            a.flags = qmm.MM_EXPORT
    so = (C.c_int32 * len(nodes))()
    rc = lib.b200_plan_analyze(args, len(nodes), C.byref(ps), so)
    arena = lib.b200_plan_arena_bytes(args, len(nodes), C.byref(ps))
    ok = rc == qmm.OK and list(so) == [src for _, _, src in nodes] and arena == (sum((m + 15) // 16 * 16 for m, _, _ in nodes) + sum(((k + k // 32 * 8) // 8 + 15) // 16 * 16 for _, k, _ in nodes)) * 8
    everyone = [None] * world
    dist.all_gather_object(everyone, (list(so), arena, [(s.r0, s.r1) for s in splits]))
    ok = ok and all(e[0] == everyone[0][0] and e[1] == everyone[0][1] for e in everyone)
    for i, (m, _, _) in enumerate(nodes):                             # the slices tile every matrix exactly once
        covered = [r for e in everyone for r in range(*e[2][i])]
        ok = ok and covered == list(range(m))
    # play the plan: the dependency of op i is so[i]; slices are exchanged per op
    full = []
    for i, (m, k, _) in enumerate(nodes):
        cur = x if so[i] < 0 else full[so[i]]
        sp = splits[i]
        mine = torch.zeros(sp.chunk)
        if sp.rows > 0:
            y = oracle.mul_mat(Q4_0, np.ascontiguousarray(wires[i][sp.r0:sp.r1]), k, sp.rows, 1, 1, cur.reshape(1, 1, 1, k))[0, 0, 0]
            mine[:sp.rows] = torch.from_numpy(y)
        gathered = torch.zeros(world * sp.chunk)
        dist.all_gather_into_tensor(gathered, mine)
        full.append(gathered[:m].numpy().copy())
    ref = []
    for i, (m, k, src) in enumerate(nodes):
        cur = x if src < 0 else ref[src]
        ref.append(oracle.mul_mat(Q4_0, wires[i], k, m, 1, 1, cur.reshape(1, 1, 1, k))[0, 0, 0])
    ok = ok and all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(full, ref))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_split_plan_host_logic_gloo(world):
    # [fc_in, v, q, k <- x; o <- v; fc_out <- fc_in] + head at reduced width; 1000 and 257 do not divide evenly
    E, F = 256, 1024
    nodes = [(F, E, -1), (E, E, -1), (E, E, -1), (E, E, -1), (E, E, 1), (E, F, 0), (257, E, 5), (1000, E, 5)]
    port = 29900 + (os.getpid() % 300) + world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_plan_worker, args=(world, port, nodes, ret), nprocs=world, join=True)
    assert all(ret.get(r, False) for r in range(world)), dict(ret)
