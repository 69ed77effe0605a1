"""Row-split tensor parallelism of one quantized mul_mat across the ranks of a torch.distributed group.

What it stands in for: the reference's split buffer + ggml_cuda_op_mul_mat multi-device loop
(src/ggml-cuda.cu:578-975, :1360-1647): rows of src0 are divided across devices, src1 is replicated, every
device computes its dst row slice.  Differences by design: one process per GPU, the dst slices are re-assembled
on EVERY rank with one all-gather over NVLink (torch.distributed / NCCL) instead of peer memcpys to a main GPU,
and each rank quantizes src1 itself (cheaper than shipping Q8_0 around).

Layout facts used here (dst is [n][m] with m contiguous, src/ggml.c:4834-4838):
  * rank r owns rows [r*c, min((r+1)*c, m)) with c = ceil(m / world): equal chunks, only the last may be short;
  * n == 1: the slices are contiguous in dst, so the all-gather runs in place on dst (padded to world*c floats);
  * n  > 1: a rank's slice dst[:, r0:r1] is strided; it is computed into a dense [n][c] staging block, all-gathered
    into [world][n][c] and permuted into dst with one strided copy (the issue the reference notes at
    src/ggml-cuda.cu:1592-1608).
torch is plumbing only (buffers, the collective, the permute copy); the mul_mat itself is `compute_slice`.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class RowSplit:
    m: int
    world: int
    rank: int

    @property
    def chunk(self) -> int:
        return (self.m + self.world - 1) // self.world

    @property
    def r0(self) -> int:
        return min(self.rank * self.chunk, self.m)

    @property
    def r1(self) -> int:
        return min((self.rank + 1) * self.chunk, self.m)

    @property
    def rows(self) -> int:
        return self.r1 - self.r0

    @property
    def padded_m(self) -> int:
        return self.chunk * self.world

    def ranges(self):
        c = self.chunk
        return [(min(r * c, self.m), min((r + 1) * c, self.m)) for r in range(self.world)]


def message_bytes(m: int, n: int, world: int) -> int:
    """bytes each rank contributes to the all-gather of one mul_mat"""
    return RowSplit(m, world, 0).chunk * n * 4


def gathered_mul_mat(dist, split: RowSplit, n: int, compute_slice, dst_full, staging=None, gathered=None, group=None):
    """dst_full: torch float32 tensor with >= n * padded_m elements (n == 1) or >= n * m (n > 1), on the rank's device.
    compute_slice(out_tensor, ld) must write this rank's rows as out[j * ld + i] for column j, local row i.
    Returns a [n, m] view of the assembled result (valid on every rank)."""
    c, m, world, rank = split.chunk, split.m, split.world, split.rank
    if n == 1:
        full = dst_full[: c * world]
        mine = full[rank * c:(rank + 1) * c]
        if split.rows > 0:
            compute_slice(mine, c)
        if world > 1:
            dist.all_gather_into_tensor(full, mine, group=group)
        return full[:m].view(1, m)
    assert staging is not None and gathered is not None
    st = staging[: n * c].view(n, c)
    if split.rows > 0:
        compute_slice(st, c)
    g = gathered[: world * n * c]
    if world > 1:
        dist.all_gather_into_tensor(g, st.reshape(-1), group=group)
    else:
        g.copy_(st.reshape(-1))
    out = dst_full[: n * m].view(n, m)
    gv = g.view(world, n, c)
    for r, (a, b) in enumerate(split.ranges()):      # world small strided copies: [n, rows_r] each
        if b > a:
            out[:, a:b].copy_(gv[r, :, : b - a])
    return out


def plan_split(plan_split_cls, splits, world: int, rank: int):
    """The b200_plan_split (include/ggml_b200.h) of a row-split decode plan for this rank: op i is this rank's slice
    splits[i] (a RowSplit) of a matrix with splits[i].m rows.  peer_arena[] is left for the caller (IPC handles)."""
    import ctypes as C
    ps = plan_split_cls()
    ps.world, ps.rank = world, rank
    row0 = (C.c_int64 * len(splits))(*[s.r0 for s in splits])
    mtot = (C.c_int64 * len(splits))(*[s.m for s in splits])
    ps.row0, ps.m_total = row0, mtot
    ps._keepalive = (row0, mtot)        # the struct only holds pointers
    return ps
