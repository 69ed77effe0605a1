"""Decode plans (b200_plan_*): a dependent sequence of decode mul_mats as one persistent launch, against the oracle
(ggml_compute_forward_mul_mat, src/ggml.c:11808, applied node by node) and against the node-by-node CUDA path
(bit-identical by construction)."""
import ctypes as C

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


PLAN_OPTION_DEFAULTS = {"plan_pub_min_k": 4096, "plan_pub_dist": 2, "plan_l2_window": 8, "plan_evict_first": 1, "plan_slots": 0}


@pytest.fixture
def plan_options(gpu_ctx):
    """set decode-plan options on the shared context for one test, defaults restored afterwards"""
    def set_(**kw):
        for k, v in kw.items():
            gpu_ctx.set_option(k, v)
    yield set_
    for k, v in PLAN_OPTION_DEFAULTS.items():
        gpu_ctx.set_option(k, v)


@pytest.mark.parametrize("dist", [1, 2])
@pytest.mark.parametrize("case", [(Q4_0, "block_pair"), (Q8_0, "block_pair"), (Q4_0, "ragged_chain"), (Q4_0, "many_ops"), (Q4_0, "k_split_8"),
                                  (Q8_0, "k_split_8")], ids=lambda c: f"{c[1]}-{c[0]}")
def test_plan_with_every_possible_src1_published(gpu_ctx, oracle, qmm, plan_options, case, dist):
    """plan_pub_min_k = 256: EVERY in-plan src1 that is not shared with the previous op is quantized once per GPU by the
    publisher warps (every CTA 1/grid of the blocks, plain activation planes + an arrival counter) and brought into a rotating
    activation buffer by the fetcher's bulk copy.  dist = 1 also publishes right behind the producing op (the publisher then
    waits for it).  quantize_row_q8_0 on the same fp32 values: the same bits as ever."""
    plan_options(plan_pub_min_k=256, plan_pub_dist=dist)
    check_dag(gpu_ctx, oracle, qmm, case[0], case[1])


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("dag", ["block_pair", "k_split_8"])
def test_plan_with_every_src1_quantized_per_cta(gpu_ctx, oracle, qmm, plan_options, qtype, dag):
    """plan_pub_min_k = 0: no publisher / fetcher work at all; every CTA fetches and quantizes every vector itself."""
    plan_options(plan_pub_min_k=0)
    check_dag(gpu_ctx, oracle, qmm, qtype, dag)


@pytest.mark.parametrize("window", [0, 2, 64])
@pytest.mark.parametrize("slots", [0, 3])
def test_plan_l2_prefetch_window_and_short_ring(gpu_ctx, oracle, qmm, plan_options, window, slots):
    """the L2 prefetcher off / barely ahead / far ahead of the weight stream, with the full ring and with three slots
    (every hand-off then stalls the producer): prefetching never changes a bit"""
    plan_options(plan_l2_window=window, plan_slots=slots)
    check_dag(gpu_ctx, oracle, qmm, Q4_0, "block_pair")
    check_dag(gpu_ctx, oracle, qmm, Q8_0, "ragged_chain")


def test_plan_accepts_allocator_style_buffer_reuse(gpu_ctx, oracle, qmm):
    """ggml_gallocr hands the memory of a dead intermediate to a later tensor (src/ggml-alloc.c).  A chain x -> a -> b -> c -> d
    with c in a's memory and d in b's: the plan keeps a and b out of plain memory (their values travel as tagged vectors) and
    c, d come out right."""
    nodes = [(1024, 512, -1), (768, 1024, 0), (1024, 768, 1), (768, 1024, 2), (300, 768, 3)]
    ws = build_dag(oracle, qmm, gpu_ctx, Q4_0, nodes, seed=77)
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, 512).astype(np.float32)
    xd = gpu_ctx.to_device(x)
    bufA, bufB, bufE = gpu_ctx.alloc(1024 * 4), gpu_ctx.alloc(768 * 4), gpu_ctx.alloc(300 * 4)
    dst = [bufA, bufB, bufA, bufB, bufE]
    args = [gpu_ctx.make_args(ws[i][0], xd.ptr if s < 0 else dst[s].ptr, 1, dst[i].ptr) for i, (m, k, s) in enumerate(nodes)]
    arr = (qmm.MulMatArgs * len(args))(*args)
    so = (C.c_int32 * len(args))()
    assert gpu_ctx.lib.b200_plan_analyze(arr, len(args), None, so) == qmm.OK and list(so) == [-1, 0, 1, 2, 3]
    plain = (C.c_int32 * len(args))()
    assert gpu_ctx.lib.b200_plan_plain_stores(arr, len(args), None, plain) == qmm.OK and list(plain) == [0, 0, 1, 1, 1]
    plan = gpu_ctx.plan_create(args)
    try:
        for _ in range(3):
            gpu_ctx.plan_launch(plan)
        gpu_ctx.synchronize()
        got_c, got_d, got_e = bufA.download(np.float32, 1024), bufB.download(np.float32, 768), bufE.download(np.float32, 300)
        # sequential reference through b200_mul_mat (which is what ggml would run node by node on this very memory)
        for a in args:
            gpu_ctx._check(gpu_ctx.lib.b200_mul_mat(gpu_ctx.h, a))
        gpu_ctx.synchronize()
        assert np.array_equal(got_c, bufA.download(np.float32, 1024))
        assert np.array_equal(got_d, bufB.download(np.float32, 768))
        assert np.array_equal(got_e, bufE.download(np.float32, 300))
        ref = run_oracle(oracle, Q4_0, nodes, ws, x)
        assert nmse(got_e, ref[4]) <= 1e-4       # (a five-deep chain: the oracle's own inputs differ by summation order)
    finally:
        gpu_ctx.plan_destroy(plan)
        xd.free(); bufA.free(); bufB.free(); bufE.free()
        for t, _ in ws:
            t.free()


def test_plan_wait_is_bounded(gpu_ctx, oracle, qmm, plan_options):
    """A row-split plan whose peer never shows up: the kernel must give up after plan_timeout_ms and b200_synchronize must
    say so, instead of hanging the GPU (world = 2, both 'peer' arenas local, rank 1 never launched)."""
    plan_options(plan_timeout_ms=300)
    nodes = [(256, 256, -1), (256, 512, 0)]          # rank 0 owns rows 0..255 of a 512-row op 0; op 1 needs all 512
    ws = build_dag(oracle, qmm, gpu_ctx, Q4_0, nodes, seed=3)
    xd = gpu_ctx.to_device(np.ones(256, np.float32))
    outs = [gpu_ctx.alloc(512 * 4), gpu_ctx.alloc(256 * 4)]
    args = [gpu_ctx.make_args(ws[0][0], xd.ptr, 1, outs[0].ptr), gpu_ctx.make_args(ws[1][0], outs[0].ptr, 1, outs[1].ptr)]
    args[1].flags |= qmm.MM_EXPORT
    split = qmm.PlanSplit()
    split.world, split.rank = 2, 0
    row0 = (C.c_int64 * 2)(0, 0)
    mtot = (C.c_int64 * 2)(512, 256)
    split.row0, split.m_total = row0, mtot
    nbytes = gpu_ctx.plan_arena_bytes(args, split)
    arena0, arena1 = gpu_ctx.alloc(nbytes), gpu_ctx.alloc(nbytes)
    for a in (arena0, arena1):
        gpu_ctx._check(gpu_ctx.lib.b200_memset(gpu_ctx.h, a.ptr, 0, nbytes))
    split.peer_arena[0], split.peer_arena[1] = arena0.ptr, arena1.ptr
    plan = gpu_ctx.plan_create(args, split)
    try:
        gpu_ctx.plan_launch(plan)
        with pytest.raises(qmm.B200Error) as e:
            gpu_ctx.synchronize()
        assert e.value.code == qmm.ERR_CUDA and "gave up waiting" in str(e.value)
        gpu_ctx.synchronize()            # the context stays usable
    finally:
        gpu_ctx.set_option("plan_timeout_ms", 0)
        gpu_ctx.plan_destroy(plan)
        xd.free(); arena0.free(); arena1.free()
        for o in outs:
            o.free()
        for t, _ in ws:
            t.free()


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
        a1 = gpu_ctx.make_args(ws[1][0], x.ptr, 1, x.ptr)         # writes over the outside vector other CTAs may still have to read
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
