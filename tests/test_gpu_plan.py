"""Decode plans (b200_plan_*): a dependent sequence of decode mul_mats as one persistent launch, against the oracle
(ggml_compute_forward_mul_mat, src/ggml.c:11808, applied node by node) and against the node-by-node CUDA path
(bit-identical by construction)."""
import os

import numpy as np
import pytest

from conftest import Q4_0, Q8_0, MUL_MAT_NMSE_TOL, nmse

pytestmark = pytest.mark.gpu


def build_dag(oracle, qmm, ctx, qtype, nodes, seed):
    """nodes: [(m, k, src)] with src = producing node or -1 (the outside vector)."""
    rng = np.random.default_rng(seed)
    ws = []
    for i, (m, k, _) in enumerate(nodes):
        w = rng.uniform(-1, 1, (m, k)).astype(np.float32) * np.float32(np.sqrt(3.0 / k))      # unit gain: deep chains stay O(1)
        wire = oracle.quantize_weights(qtype, w)
        t = qmm.QTensor(ctx, qtype, k, m)
        t.set(wire)
        ws.append((t, wire))
    return ws


def run_oracle(oracle, qtype, nodes, ws, x, inputs=None):
    """node i = W_i x src.  inputs = the vectors the device actually fed each node (per-node parity: a deep chain amplifies
    the ~1e-9 summation-order differences through its quantization steps, which is not what is being tested); None = chain
    the oracle's own outputs."""
    outs = []
    for i, (m, k, src) in enumerate(nodes):
        cur = x if src < 0 else (outs[src] if inputs is None else inputs[src])
        outs.append(oracle.mul_mat(qtype, ws[i][1], k, m, 1, 1, cur.reshape(1, 1, 1, k))[0, 0, 0])
    return outs


DAGS = {
    # a GPT-J-like block pair at reduced width: q,k,v,fc_in <- x; o <- v; fc_out <- fc_in (k-split, G = 4); next block <- fc_out
    "block_pair": [(512, 512, -1), (512, 512, -1), (512, 512, -1), (16384, 512, -1), (512, 512, 2), (512, 16384, 3),
                   (512, 512, 5), (768, 512, 5), (1000, 512, 5), (8192, 512, 5), (512, 768, 7), (301, 8192, 9)],
    # fewer rows than CTAs, odd row counts, a chain
    "ragged_chain": [(1024, 256, -1), (7, 1024, 0), (256, 256, -1), (2049, 256, 2), (33, 256, 2), (4352, 256, 2), (5, 4352, 5)],
    "single": [(4096, 4096, -1)],
    # many same-input ops: no barrier between them, warps free-run around the ring for many laps
    "free_run": [(4096, 4096, -1)] * 6 + [(2048, 4096, -1)] * 6 + [(1000, 4096, 0)] * 4,
    # more ops than one shared-memory descriptor window (128) holds; op numbers above 127 in the tags
    "many_ops": [(256, 256, -1)] + [(256, 256, i) for i in range(139)],
    # k split over 8 warps (k > 16384): Llama-2-70B-like up / down at reduced m, and the largest k the path takes
    "k_split_8": [(28672, 512, -1), (512, 28672, 0), (32768, 512, -1), (100, 32768, 2)],
}


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("dag", sorted(DAGS))
def test_plan_matches_oracle_and_node_by_node(gpu_ctx, oracle, qmm, qtype, dag):
    check_dag(gpu_ctx, oracle, qmm, qtype, dag)


@pytest.mark.parametrize("case", [(Q4_0, "block_pair"), (Q8_0, "block_pair"), (Q4_0, "ragged_chain"), (Q4_0, "many_ops"), (Q8_0, "k_split_8")],
                         ids=lambda c: f"{c[1]}-{c[0]}")
@pytest.mark.parametrize("ring", ["-2", "-1", "3"])
def test_plan_with_ring_fed_src1(gpu_ctx, oracle, qmm, monkeypatch, case, ring):
    """B200_PLAN_LL_RING: src1 vectors produced inside the plan travel through the weight ring (copied by the producer
    thread) instead of being fetched from L2 by the consumers.  -2 feeds EVERY such vector that way, also those whose
    producer is the op right before (the copy is then always taken too early: the tag check must catch it and fall back),
    -1 those produced two or more ops back, 3 the production rule (at least 3 ring slots in between).  Same bits as ever."""
    monkeypatch.setenv("B200_PLAN_LL_RING", ring)
    check_dag(gpu_ctx, oracle, qmm, case[0], case[1])


@pytest.mark.parametrize("case", [(Q4_0, "block_pair"), (Q8_0, "block_pair"), (Q4_0, "ragged_chain"), (Q4_0, "many_ops"), (Q4_0, "k_split_8"),
                                  (Q8_0, "k_split_8")], ids=lambda c: f"{c[1]}-{c[0]}")
def test_plan_with_src1_quantized_once_per_gpu(gpu_ctx, oracle, qmm, monkeypatch, case):
    """B200_PLAN_LLQ=k_min: an in-plan src1 of k >= k_min is quantized once per GPU (every CTA 1/grid of the blocks, published
    as tagged words) instead of once per CTA.  quantize_row_q8_0 on the same fp32 values: the same bits as ever."""
    monkeypatch.setenv("B200_PLAN_LLQ", "256")
    monkeypatch.setenv("B200_PLAN_LLQ_DIST", "1")       # also right behind the producing op (the publisher then waits for it)
    check_dag(gpu_ctx, oracle, qmm, case[0], case[1])


@pytest.mark.skipif(os.environ.get("B200_TEST_EXPERIMENTAL") != "1", reason="unvalidated kernel mode: set B200_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("dist", ["1", "2"])
@pytest.mark.parametrize("case", [(Q4_0, "block_pair"), (Q8_0, "block_pair"), (Q4_0, "ragged_chain"), (Q4_0, "many_ops"), (Q4_0, "k_split_8"),
                                  (Q8_0, "k_split_8")], ids=lambda c: f"{c[1]}-{c[0]}")
def test_plan_with_published_planes_experimental(gpu_ctx, oracle, qmm, monkeypatch, case, dist):
    """B200_PLAN_PUBQ=1 (kernel MODE 8, written at the end of round 1: only the block_pair cases have run on a GPU so far -- they pass --
    and no timing exists, hence the gate): the once-per-GPU quantization
    published as plain activation planes + an arrival counter, taken by every CTA with one bulk copy.  Same bits as ever."""
    monkeypatch.setenv("B200_PLAN_PUBQ", "1")
    monkeypatch.setenv("B200_PLAN_LLQ", "256")
    monkeypatch.setenv("B200_PLAN_LLQ_DIST", dist)
    check_dag(gpu_ctx, oracle, qmm, case[0], case[1])


@pytest.mark.skipif(os.environ.get("B200_TEST_EXPERIMENTAL") != "1", reason="unvalidated kernel mode: set B200_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("llq", ["0", "256"])
@pytest.mark.parametrize("case", [(Q4_0, "block_pair"), (Q8_0, "block_pair"), (Q4_0, "k_split_8"), (Q8_0, "k_split_8"), (Q4_0, "ragged_chain")],
                         ids=lambda c: f"{c[1]}-{c[0]}")
def test_plan_without_k_split_teams_experimental(gpu_ctx, oracle, qmm, monkeypatch, case, llq):
    """B200_PLAN_NOSPLIT=1 (kernel MODE bit 16, never run): ops with k > 4096 are walked segment by segment by the slot's one
    warp instead of being split over a team of warps with a partials pass.  Partials are added in the same order: same bits
    (except the Q8_0 k = 32768 node, which b200_mul_mat itself serves with another kernel: NMSE there, as in check_dag)."""
    monkeypatch.setenv("B200_PLAN_NOSPLIT", "1")
    monkeypatch.setenv("B200_PLAN_LLQ", llq)
    check_dag(gpu_ctx, oracle, qmm, case[0], case[1])


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
def test_plan_with_every_src1_quantized_per_cta(gpu_ctx, oracle, qmm, monkeypatch, qtype):
    """B200_PLAN_LLQ=0: the kernel without the publisher warp (what row-split plans run), on the DAG whose k = 16384 node takes
    the quantized-once path by default."""
    monkeypatch.setenv("B200_PLAN_LLQ", "0")
    check_dag(gpu_ctx, oracle, qmm, qtype, "block_pair")


def check_dag(gpu_ctx, oracle, qmm, qtype, dag):
    nodes = DAGS[dag]
    ws = build_dag(oracle, qmm, gpu_ctx, qtype, nodes, seed=len(nodes) * 7 + qtype)
    rng = np.random.default_rng(99)
    k0 = next(k for (_, k, s) in nodes if s < 0)
    xs = {}
    for (_, k, s) in nodes:
        if s < 0 and k not in xs:
            xs[k] = rng.uniform(-1, 1, k).astype(np.float32)
    assert len(xs) == 1, "test DAGs use one outside vector"
    x = xs[k0]
    xd = gpu_ctx.to_device(x)
    outs = [gpu_ctx.alloc(m * 4) for (m, _, _) in nodes]
    for o in outs:
        gpu_ctx._check(gpu_ctx.lib.b200_memset(gpu_ctx.h, o.ptr, 0xff, o.nbytes))
    args = [gpu_ctx.make_args(ws[i][0], xd.ptr if s < 0 else outs[s].ptr, 1, outs[i].ptr) for i, (m, k, s) in enumerate(nodes)]
    plan = gpu_ctx.plan_create(args)
    try:
        for rep in range(3):          # replays: tags must stay unique from launch to launch
            gpu_ctx.plan_launch(plan)
        gpu_ctx.synchronize()
        got = [outs[i].download(np.float32, nodes[i][0]) for i in range(len(nodes))]
        ref = run_oracle(oracle, qtype, nodes, ws, x, inputs=got)
        for i in range(len(nodes)):
            assert np.isfinite(got[i]).all(), f"node {i}"
            assert nmse(got[i], ref[i]) <= MUL_MAT_NMSE_TOL, f"node {i}: nmse {nmse(got[i], ref[i])}"
        # node by node through b200_mul_mat: same arithmetic, same summation order -> same bits
        for o in outs:
            gpu_ctx._check(gpu_ctx.lib.b200_memset(gpu_ctx.h, o.ptr, 0, o.nbytes))
        for a in args:
            gpu_ctx._check(gpu_ctx.lib.b200_mul_mat(gpu_ctx.h, a))
        gpu_ctx.synchronize()
        # (b200_mul_mat serves Q8_0 rows of k = 32768 with the generic GEMV, whose summation order differs: NMSE there)
        exact = not (qtype == Q8_0 and dag == "k_split_8")
        for i in range(len(nodes)):
            one = outs[i].download(np.float32, nodes[i][0])
            if exact:
                assert np.array_equal(got[i], one), f"node {i} differs from b200_mul_mat"
            else:
                assert nmse(got[i], one) <= 1e-9, f"node {i}: nmse vs b200_mul_mat {nmse(got[i], one)}"
        # a new input through the same plan
        x2 = rng.uniform(-2, 2, k0).astype(np.float32)
        xd.upload(x2)
        gpu_ctx.plan_launch(plan)
        gpu_ctx.synchronize()
        got2 = [outs[i].download(np.float32, nodes[i][0]) for i in range(len(nodes))]
        ref2 = run_oracle(oracle, qtype, nodes, ws, x2, inputs=got2)
        for i in (0, len(nodes) // 2, len(nodes) - 1):
            assert nmse(got2[i], ref2[i]) <= MUL_MAT_NMSE_TOL, f"second input, node {i}"
    finally:
        gpu_ctx.plan_destroy(plan)
        xd.free()
        for o in outs:
            o.free()
        for t, _ in ws:
            t.free()


def test_plan_rejects_what_it_cannot_run(gpu_ctx, oracle, qmm):
    ws = build_dag(oracle, qmm, gpu_ctx, Q4_0, [(64, 256, -1), (64, 256, -1)], seed=1)
    x = gpu_ctx.alloc(256 * 4)
    y = gpu_ctx.alloc(64 * 4)
    try:
        a0 = gpu_ctx.make_args(ws[0][0], x.ptr, 1, y.ptr)
        a1 = gpu_ctx.make_args(ws[1][0], x.ptr, 1, y.ptr)         # two nodes writing the same plain vector
        with pytest.raises(qmm.B200Error) as e:
            gpu_ctx.plan_create([a0, a1])
        assert e.value.code == qmm.ERR_UNSUPPORTED
        a2 = gpu_ctx.make_args(ws[1][0], x.ptr, 4, y.ptr)         # not a decode shape
        with pytest.raises(qmm.B200Error) as e:
            gpu_ctx.plan_create([a2])
        assert e.value.code == qmm.ERR_UNSUPPORTED
    finally:
        x.free(); y.free()
        for t, _ in ws:
            t.free()


def test_plan_c5_full_size_llama70b_ffn(gpu_ctx, oracle, qmm):
    """BASELINE.json C5 at full size: gate and up 28672 x 8192 read x, down 8192 x 28672 reads up (the gating product is glue
    outside this path).  Bitwise against node-by-node b200_mul_mat, sampled rows against the oracle."""
    E, F = 8192, 28672
    shapes = [(F, E, -1), (F, E, -1), (E, F, 1)]
    rng = np.random.default_rng(5)
    ws, wires = [], []
    for i, (m, k, _) in enumerate(shapes):
        wire = qmm.random_wire_weights(Q4_0, k, m, seed=50 + i)
        t = qmm.QTensor(gpu_ctx, Q4_0, k, m)
        t.set(wire)
        ws.append(t); wires.append(wire)
    x = rng.uniform(-1, 1, E).astype(np.float32)
    xd = gpu_ctx.to_device(x)
    outs = [gpu_ctx.alloc(m * 4) for (m, _, _) in shapes]
    args = [gpu_ctx.make_args(ws[i], xd.ptr if s < 0 else outs[s].ptr, 1, outs[i].ptr) for i, (m, k, s) in enumerate(shapes)]
    plan = gpu_ctx.plan_create(args)
    try:
        gpu_ctx.plan_launch(plan)
        gpu_ctx.plan_launch(plan)
        gpu_ctx.synchronize()
        got = [outs[i].download(np.float32, shapes[i][0]) for i in range(3)]
        for o in outs:
            gpu_ctx._check(gpu_ctx.lib.b200_memset(gpu_ctx.h, o.ptr, 0, o.nbytes))
        for a in args:
            gpu_ctx._check(gpu_ctx.lib.b200_mul_mat(gpu_ctx.h, a))
        gpu_ctx.synchronize()
        for i in range(3):
            assert np.array_equal(got[i], outs[i].download(np.float32, shapes[i][0])), f"node {i} differs from b200_mul_mat"
        # sampled rows of every node against the oracle on the same inputs
        for i, (m, k, s) in enumerate(shapes):
            rows = np.unique(np.concatenate([[0, 1, m - 1], rng.integers(0, m, 29)]))
            src = x if s < 0 else got[s]
            ref = oracle.mul_mat(Q4_0, np.ascontiguousarray(wires[i][rows]), k, len(rows), 1, 1, src.reshape(1, 1, 1, k))[0, 0, 0]
            assert nmse(got[i][rows], ref) <= MUL_MAT_NMSE_TOL, f"node {i}: nmse {nmse(got[i][rows], ref)}"
    finally:
        gpu_ctx.plan_destroy(plan)
        xd.free()
        for o in outs:
            o.free()
        for t in ws:
            t.free()
