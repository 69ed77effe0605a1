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
