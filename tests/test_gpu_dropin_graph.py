"""The reference host framework with this backend dropped in, on a multi-node graph: the SAME ggml graph of MUL_MAT nodes
(built with ggml_mul_mat / ggml_build_forward_expand, allocated with ggml_backend_alloc_ctx_tensors, computed with
ggml_backend_graph_compute -- oracle/ref_shim.c) runs once on the reference CPU backend (oracle/_ref/libref_shim.so) and once
on the B200 backend found through the reference's registry (oracle/_ref/libdropin_shim.so = the unmodified core compiled
with -DGGML_USE_CUDA, linked against libggml-b200-backend.so).  Decode-shaped graphs must go down as ONE persistent launch
(decode plan, ggml-b200.c:b200_try_graph_as_plan) and every node must match the CPU backend within test-backend-ops'
NMSE <= 5e-4 (tests/test-backend-ops.cpp:921-923)."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT, Q4_0, Q8_0, WIRE, MUL_MAT_NMSE_TOL, nmse

pytestmark = pytest.mark.gpu
REF = ROOT / "oracle" / "_ref"
vp = C.c_void_p


def load(name):
    path = REF / name
    assert path.exists(), f"{path} must be prebuilt (make -C oracle ref dropin) and travel with the snapshot"
    lib = C.CDLL(str(path))
    lib.ref_dag_create.restype = vp
    lib.ref_dag_create_galloc.restype = vp
    lib.ref_chain_aliased_nodes.argtypes = [vp]
    lib.ref_chain_get_out.argtypes = [vp, vp]
    lib.ref_chain_out_elements.restype = C.c_int64
    lib.ref_chain_out_elements.argtypes = [vp]
    lib.ref_chain_compute.restype = C.c_double
    lib.ref_chain_compute.argtypes = [vp]
    lib.ref_chain_compute_planned.restype = C.c_double
    lib.ref_chain_compute_planned.argtypes = [vp]
    lib.ref_chain_node_elements.restype = C.c_int64
    lib.ref_chain_node_elements.argtypes = [vp, C.c_int]
    lib.ref_chain_n_nodes.argtypes = [vp]
    lib.ref_chain_get_node.argtypes = [vp, C.c_int, vp]
    lib.ref_chain_set_weight.argtypes = [vp, C.c_int, vp]
    lib.ref_chain_set_x.argtypes = [vp, vp]
    lib.ref_chain_free.argtypes = [vp]
    lib.ref_select_backend.argtypes = [C.c_char_p]
    lib.ref_time_init()
    return lib


def make_graph(lib, qtype, nodes, shapes, ncols, threads=4, galloc=False):
    """nodes: [(weight id, src node or -1)], shapes: [(k, m)] per weight id"""
    n = len(nodes)
    h = (lib.ref_dag_create_galloc if galloc else lib.ref_dag_create)(qtype, n, (C.c_int * n)(*[w for w, _ in nodes]), (C.c_int * n)(*[s for _, s in nodes]), len(shapes),
                           (C.c_int64 * len(shapes))(*[k for k, _ in shapes]), (C.c_int64 * len(shapes))(*[m for _, m in shapes]),
                           C.c_int64(ncols), threads)
    assert h, "ref_dag_create failed (backend not in the registry?)"
    return vp(h)


def node_outputs(lib, h):
    outs = []
    for i in range(lib.ref_chain_n_nodes(h)):
        o = np.zeros(lib.ref_chain_node_elements(h, i), np.float32)
        lib.ref_chain_get_node(h, i, o.ctypes.data_as(vp))
        outs.append(o)
    return outs


# two GPT-J-like blocks at reduced width + head: [fc_in, v, q, k <- x; o <- v; fc_out <- fc_in] x 2, lm_head
E, F, V = 512, 2048, 1000
SHAPES = [(E, F), (E, E), (E, E), (E, E), (E, E), (F, E), (E, V)]     # (k, m) per weight id
NODES = [(0, -1), (1, -1), (2, -1), (3, -1), (4, 1), (5, 0),
         (0, 5), (1, 5), (2, 5), (3, 5), (4, 7), (5, 6), (6, 11)]


@pytest.mark.parametrize("planned", [False, True])
@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("ncols", [1, 3])
def test_same_ggml_graph_on_cpu_backend_and_on_b200_backend(oracle, qtype, ncols, planned):
    """planned: through ggml_backend_graph_plan_create / _compute instead of ggml_backend_graph_compute"""
    cpu = load("libref_shim.so")
    gpu = load("libdropin_shim.so")
    gpu.ref_select_backend(b"B2000")
    rng = np.random.default_rng(7 + qtype + ncols)
    hc = make_graph(cpu, qtype, NODES, SHAPES, ncols)
    hg = make_graph(gpu, qtype, NODES, SHAPES, ncols)
    try:
        for j, (k, m) in enumerate(SHAPES):
            w = oracle.quantize_weights(qtype, rng.uniform(-1, 1, (m, k)).astype(np.float32) * (2.0 / np.sqrt(k)))
            assert w.nbytes == m * (k // 32) * WIRE[qtype]
            cpu.ref_chain_set_weight(hc, j, w.ctypes.data_as(vp))
            gpu.ref_chain_set_weight(hg, j, w.ctypes.data_as(vp))
        for rep in range(3):                                     # new activations through the same (cached) graph
            x = rng.uniform(-1, 1, (ncols, E)).astype(np.float32)
            cpu.ref_chain_set_x(hc, x.ctypes.data_as(vp))
            gpu.ref_chain_set_x(hg, x.ctypes.data_as(vp))
            cpu.ref_chain_compute(hc)
            (gpu.ref_chain_compute_planned if planned else gpu.ref_chain_compute)(hg)
            want, got = node_outputs(cpu, hc), node_outputs(gpu, hg)
            assert len(want) == len(got) == len(NODES)
            for i, (a, b) in enumerate(zip(got, want)):
                assert np.isfinite(a).all(), f"node {i}"
                assert nmse(a, b) <= MUL_MAT_NMSE_TOL, f"rep {rep} node {i}: nmse {nmse(a, b)}"
        gpu.ref_chain_plan_launches.restype = C.c_int64
        gpu.ref_chain_plan_launches.argtypes = [vp]
        gpu.ref_chain_kernel_launches.restype = C.c_int64
        gpu.ref_chain_kernel_launches.argtypes = [vp]
        plans, kernels = gpu.ref_chain_plan_launches(hg), gpu.ref_chain_kernel_launches(hg)
        if ncols == 1:
            assert plans == 3 and kernels == 3, f"decode graph should be one persistent launch per compute ({plans} plans, {kernels} kernels)"
        else:
            assert plans == 0 and kernels >= 3 * len(NODES), (plans, kernels)
    finally:
        cpu.ref_chain_free(hc)
        gpu.ref_chain_free(hg)


@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("ncols", [1, 3])
def test_graph_allocated_by_gallocr_runs_as_one_plan(oracle, qtype, ncols):
    """The compute nodes allocated by ggml_gallocr_alloc_graph (examples/gpt-2/main-backend.cpp:744), which hands the memory of
    dead intermediates to later nodes: the decode graph must STILL go down as one persistent launch (the plan keeps the dead
    intermediates out of plain memory) and its output must match the CPU backend's."""
    cpu = load("libref_shim.so")
    gpu = load("libdropin_shim.so")
    gpu.ref_select_backend(b"B2000")
    rng = np.random.default_rng(17 + qtype + ncols)
    hc = make_graph(cpu, qtype, NODES, SHAPES, ncols, galloc=True)
    hg = make_graph(gpu, qtype, NODES, SHAPES, ncols, galloc=True)
    try:
        assert gpu.ref_chain_aliased_nodes(hg) > 0, "the allocator was expected to reuse memory in this graph"
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
            want = np.zeros(cpu.ref_chain_out_elements(hc), np.float32)
            got = np.zeros(gpu.ref_chain_out_elements(hg), np.float32)
            cpu.ref_chain_get_out(hc, want.ctypes.data_as(vp))
            gpu.ref_chain_get_out(hg, got.ctypes.data_as(vp))
            # (the whole chain on both sides: a few quantization stages deep, summation order differs -> the bar, not bit equality)
            assert np.isfinite(got).all() and nmse(got, want) <= MUL_MAT_NMSE_TOL, f"rep {rep}: nmse {nmse(got, want)}"
        gpu.ref_chain_plan_launches.restype = C.c_int64
        gpu.ref_chain_plan_launches.argtypes = [vp]
        gpu.ref_chain_kernel_launches.restype = C.c_int64
        gpu.ref_chain_kernel_launches.argtypes = [vp]
        plans, kernels = gpu.ref_chain_plan_launches(hg), gpu.ref_chain_kernel_launches(hg)
        if ncols == 1:
            assert plans == 3 and kernels == 3, f"the aliased decode graph should be one persistent launch per compute ({plans} plans, {kernels} kernels)"
        else:
            assert plans == 0
    finally:
        cpu.ref_chain_free(hc)
        gpu.ref_chain_free(hg)
