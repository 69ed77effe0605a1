"""Host-only logic of decode plans (b200_plan_analyze, the first half of b200_plan_create): which op reads which op's result
(ggml dataflow by tensor address, as ggml_backend_graph_compute would execute the nodes in order: src/ggml-backend.c:275-279),
and which graphs must be refused because dataflow execution would break what sequential execution hides (vectors sharing
memory, the way a graph allocator reuses buffers: src/ggml-alloc.c).  No device needed: pointers are made-up addresses."""
import ctypes as C

import numpy as np
import pytest

from conftest import Q4_0, Q8_0


def args_for(qmm, nodes, base=0x10000000, x_addr=0x0F000000, qtype=Q4_0, ncols=1):
    """nodes: [(m, k, src, dst_addr or None)] -> MulMatArgs[]; dst default: its own 1 MB slab"""
    out = (qmm.MulMatArgs * len(nodes))()
    dst = [n[3] if len(n) > 3 and n[3] is not None else base + i * 0x100000 for i, n in enumerate(nodes)]
    for i, n in enumerate(nodes):
        m, k, src = n[0], n[1], n[2]
        a = out[i]
        a.type = qtype
        a.src0_dev = 0x40000000 + i * 0x1000000
        a.src0_nblocks_total = m * (k // 32)
        a.ne00, a.ne01, a.ne02, a.ne03 = k, m, 1, 1
        a.src1_dev = x_addr if src < 0 else dst[src]
        a.ne11, a.ne12, a.ne13 = ncols, 1, 1
        a.nb11 = k * 4
        a.nb12 = a.nb13 = k * 4 * ncols
        a.dst_dev = dst[i]
    return out


def analyze(qmm, args, split=None):
    lib = qmm.load_library()
    so = (C.c_int32 * len(args))()
    rc = lib.b200_plan_analyze(args, len(args), C.byref(split) if split is not None else None, so)
    return rc, list(so)


def test_dataflow_follows_tensor_addresses(qmm):
    # fc_in, v, q, k <- x; o <- v; fc_out <- fc_in; next block <- fc_out
    nodes = [(16384, 4096, -1), (4096, 4096, -1), (4096, 4096, -1), (4096, 4096, -1), (4096, 4096, 1), (4096, 16384, 0),
             (16384, 4096, 5), (50400, 4096, 5)]
    rc, so = analyze(qmm, args_for(qmm, nodes))
    assert rc == qmm.OK
    assert so == [-1, -1, -1, -1, 1, 0, 5, 5]


def test_latest_writer_of_an_address_is_the_producer(qmm):
    # op 2 re-reads the outside vector; ops 0 and 3 must not be confused although 3 comes later
    nodes = [(256, 256, -1), (256, 256, 0), (256, 256, -1), (512, 256, 1)]
    rc, so = analyze(qmm, args_for(qmm, nodes))
    assert rc == qmm.OK and so == [-1, 0, -1, 1]


def plain_stores(qmm, args):
    lib = qmm.load_library()
    out = (C.c_int32 * len(args))()
    rc = lib.b200_plan_plain_stores(args, len(args), None, out)
    return rc, list(out)


def test_allocator_style_buffer_reuse_drops_the_dead_store(qmm):
    """ggml_gallocr gives a later tensor the memory of a dead intermediate (src/ggml-alloc.c).  Dataflow execution has no
    global order between a slow CTA's store of the dead tensor and a fast CTA's store of the new one, so the dead store is
    dropped; its value travels as a tagged vector.  Dataflow still follows the LATEST writer of an address."""
    nodes = [(256, 256, -1, 0x20000000), (256, 256, 0), (256, 256, 1, 0x20000000), (128, 256, 2)]
    args = args_for(qmm, nodes)
    assert analyze(qmm, args) == (qmm.OK, [-1, 0, 1, 2])
    assert plain_stores(qmm, args) == (qmm.OK, [0, 1, 1, 1])
    # a smaller tensor inside the dead one's block
    nodes = [(512, 256, -1, 0x20000000), (256, 512, 0), (64, 256, 1, 0x20000000 + 1024)]
    args = args_for(qmm, nodes)
    assert analyze(qmm, args) == (qmm.OK, [-1, 0, 1])
    assert plain_stores(qmm, args) == (qmm.OK, [0, 1, 1])
    # an op may take the memory of the outside input once its own producer has consumed it everywhere
    nodes = [(256, 256, -1), (256, 256, 0, 0x0F000000)]
    assert analyze(qmm, args_for(qmm, nodes))[0] == qmm.OK


@pytest.mark.parametrize("case", ["dst_over_live_input", "dst_over_input_of_later_reader", "src_inside_dst"])
def test_hazards_dataflow_execution_cannot_order_are_refused(qmm, case):
    if case == "dst_over_live_input":      # op 1 does not wait for op 0: some CTA may still have to read x for op 0
        nodes = [(256, 256, -1), (256, 256, -1, 0x0F000000)]
    elif case == "dst_over_input_of_later_reader":   # op 2 waits only for op 0, but op 1 reads x too and may not have started
        nodes = [(256, 256, -1), (512, 256, -1), (256, 256, 0, 0x0F000000)]
    else:                            # src1 points INTO another op's dst (a view): not "exactly that vector"
        nodes = [(512, 256, -1, 0x20000000), (256, 256, -1)]
    args = args_for(qmm, nodes)
    if case == "src_inside_dst":
        args[1].src1_dev = 0x20000000 + 1024
    rc, _ = analyze(qmm, args)
    assert rc == qmm.ERR_UNSUPPORTED


@pytest.mark.parametrize("bad", ["ncols", "k_not_256", "k_too_long", "mixed_types", "batched", "force_gemm", "size_mismatch"])
def test_shapes_outside_the_decode_path_are_refused(qmm, bad):
    nodes = [(256, 512, -1), (128, 256, 0)]
    args = args_for(qmm, nodes)
    if bad == "ncols":
        args = args_for(qmm, nodes, ncols=2)
    elif bad == "k_not_256":
        args[0].ne00 = 480
    elif bad == "k_too_long":
        args[0].ne00 = 65536
        args[0].src0_nblocks_total = 256 * 2048
    elif bad == "mixed_types":
        args[1].type = Q8_0
    elif bad == "batched":
        args[1].ne02 = 2
    elif bad == "force_gemm":
        args[1].flags = qmm.MM_FORCE_GEMM
    elif bad == "size_mismatch":     # reads op 0's dst but with another length
        args[1].ne00 = 512
        args[1].src0_nblocks_total = 128 * 16
    rc, _ = analyze(qmm, args)
    assert rc == qmm.ERR_UNSUPPORTED


def test_bad_arguments_are_invalid_not_unsupported(qmm):
    lib = qmm.load_library()
    args = args_for(qmm, [(256, 256, -1)])
    assert lib.b200_plan_analyze(args, 0, None, None) == qmm.ERR_INVALID
    args[0].src0_nblocks_total = 10          # src0 does not hold m * k/32 blocks
    assert analyze(qmm, args)[0] == qmm.ERR_INVALID
    args = args_for(qmm, [(256, 256, -1)])
    args[0].src1_dev = 0
    assert analyze(qmm, args)[0] == qmm.ERR_INVALID


def test_row_split_description_is_checked(qmm):
    # rank 1 of 2 holds rows 128..255 of a 256-row matrix, rows 25200..50399 of the head
    nodes = [(128, 256, -1), (25200, 256, 0)]
    args = args_for(qmm, nodes)
    s = qmm.PlanSplit()
    s.world, s.rank = 2, 1
    row0 = (C.c_int64 * 2)(128, 25200)
    mtot = (C.c_int64 * 2)(256, 50400)
    s.row0, s.m_total = row0, mtot
    rc, so = analyze(qmm, args, s)
    assert rc == qmm.OK and so == [-1, 0]          # op 1 reads op 0's WHOLE vector (m_total == k)
    mtot[0] = 200                                  # slice does not fit the matrix
    assert analyze(qmm, args, s)[0] == qmm.ERR_INVALID
    mtot[0] = 256
    s.rank = 2
    assert analyze(qmm, args, s)[0] == qmm.ERR_INVALID


def test_arena_bytes_cover_tagged_vectors_and_published_planes(qmm):
    """b200_plan_arena_bytes is what a row-split caller allocates (and shares over CUDA IPC) before b200_plan_create: one
    tagged vector per op (8 B per element, padded to 128 B), plus room for the published planes of every op's src1
    (k int8 + k/32 x (fp32 d, 8 * sum)), whatever the options -- every rank must compute the same layout."""
    lib = qmm.load_library()
    nodes = [(16384, 4096, -1), (4096, 4096, -1), (4096, 16384, 0), (1000, 4096, 2)]
    args = args_for(qmm, nodes)
    pad16 = lambda n: (n + 15) // 16 * 16
    ll = sum(pad16(m) for m, _, _ in nodes) * 8
    pub = sum(pad16((k + k // 32 * 8) // 8) for _, k, _ in nodes) * 8
    assert lib.b200_plan_arena_bytes(args, len(nodes), None) == ll + pub
    split = qmm.PlanSplit()
    split.world, split.rank = 2, 0
    row0 = (C.c_int64 * len(nodes))(*[0] * len(nodes))
    mtot = (C.c_int64 * len(nodes))(*[2 * m for m, _, _ in nodes])
    split.row0, split.m_total = row0, mtot
    ll2 = sum(pad16(2 * m) for m, _, _ in nodes) * 8
    assert lib.b200_plan_arena_bytes(args, len(nodes), C.byref(split)) == ll2 + pub


def published(qmm, args, sm_count=148, split=None, min_k=0, dist=0):
    lib = qmm.load_library()
    out = (C.c_int32 * len(args))()
    rc = lib.b200_plan_published(args, len(args), C.byref(split) if split is not None else None, sm_count, min_k, dist, out)
    return rc, list(out)


def test_which_src1_vectors_are_quantized_once_per_gpu(qmm):
    """b200_plan_published: the host-side choice behind the publisher / fetcher warps.  GPT-J block:
    fc_in, v, q, k <- x; o <- v; fc_out <- fc_in; the next block's fc_in, v <- fc_out."""
    nodes = [(16384, 4096, -1), (4096, 4096, -1), (4096, 4096, -1), (4096, 4096, -1), (4096, 4096, 1), (4096, 16384, 0),
             (16384, 4096, 5), (4096, 4096, 5)]
    args = args_for(qmm, nodes)
    # default (min_k 4096, dist 2): o <- v and fc_out <- fc_in; fc_in <- fc_out is a true dependency (distance 1)
    assert published(qmm, args) == (qmm.OK, [0, 0, 0, 0, 1, 1, 0, 0])
    assert published(qmm, args, min_k=8192) == (qmm.OK, [0, 0, 0, 0, 0, 1, 0, 0])
    # ... unless asked for; v <- fc_out shares fc_in's input: never
    assert published(qmm, args, min_k=4096, dist=1) == (qmm.OK, [0, 0, 0, 0, 1, 1, 1, 0])
    # a device with few SMs: 512 blocks over 8 CTAs are more than one warp's run of 16 blocks
    assert published(qmm, args, sm_count=8, min_k=8192) == (qmm.OK, [0] * 8)
    assert published(qmm, args, sm_count=64, min_k=8192)[1][5] == 1
    # row-split plans: the same choice (every rank quantizes for its own arena)
    split = qmm.PlanSplit()
    split.world, split.rank = 2, 1
    nodes2 = [(m // 2, k, src) for m, k, src in nodes]
    row0 = (C.c_int64 * len(nodes))(*[m // 2 for m, _, _ in nodes])
    mtot = (C.c_int64 * len(nodes))(*[m for m, _, _ in nodes])
    split.row0, split.m_total = row0, mtot
    args2 = args_for(qmm, nodes2)
    assert published(qmm, args2, split=split) == (qmm.OK, [0, 0, 0, 0, 1, 1, 0, 0])
    # bad arguments
    assert published(qmm, args, sm_count=0)[0] == qmm.ERR_INVALID
