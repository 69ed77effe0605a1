#!/usr/bin/env python3
"""Generate tests/golden/qmm_golden.npz from the UNMODIFIED reference (oracle/_ref/libref_shim.so).

The reference ships no golden vectors for this path, so these are outputs of the
reference itself, run in the build container where /root/reference exists:
  * quantize_row_q8_0 (runtime from_float, AVX2 branch)      src/ggml-quants.c:465, :535-618
  * ggml_quantize_chunk Q4_0 / Q8_0 (weights)                 src/ggml.c:21594-21623
  * ggml_vec_dot_q4_0_q8_0 / ggml_vec_dot_q8_0_q8_0           src/ggml-quants.c:3469, :4819
  * MUL_MAT through the CPU backend for the test-backend-ops shapes (tests/test-backend-ops.cpp:2067-2098)
Run:  make -C oracle ref && python oracle/make_golden.py
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "oracle" / "_ref" / "libref_shim.so"
OUT = ROOT / "tests" / "golden" / "qmm_golden.npz"
Q4_0, Q8_0 = 2, 8
WIRE = {Q4_0: 18, Q8_0: 34}
vp = C.c_void_p


def main():
    if not SHIM.exists():
        sys.exit(f"{SHIM} missing: run `make -C oracle ref` where /root/reference exists")
    r = C.CDLL(str(SHIM))
    r.ref_vec_dot.restype = C.c_float
    r.ref_quantize_chunk.restype = C.c_size_t
    r.ref_mm_create.restype = vp
    r.ref_mm_compute.restype = C.c_double
    r.ref_time_init()
    rng = np.random.default_rng(20240518)
    g = {}

    # --- activations: a few regimes, incl. zero blocks, tiny, huge, exact .5 ties after scaling
    k = 256
    x = rng.uniform(-1, 1, (12, k)).astype(np.float32)
    x[1, :32] = 0.0
    x[2] *= 1e-30
    x[3] *= 1e30
    x[4, :32] = np.arange(32, dtype=np.float32) - 15.5          # ties: amax 16.5
    x[5, :32] = (np.arange(32, dtype=np.float32) - 16.0) / 254.0 * 127.0
    x[6] = rng.standard_normal(k).astype(np.float32) * 3
    x[7, 32:64] = 127.0
    x[8, 64:96] = -0.0
    q8 = np.zeros((x.shape[0], k // 32 * 34), np.uint8)
    for i in range(x.shape[0]):
        r.ref_from_float(Q8_0, x[i].ctypes.data_as(vp), q8[i].ctypes.data_as(vp), C.c_int64(k))
    g["act_x"], g["act_q8_0"] = x, q8

    # --- weights through ggml_quantize_chunk (imatrix == NULL -> *_reference quantizers)
    m = 24
    wf = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    wf[0, :32] = 0.0
    for t, name in ((Q4_0, "q4_0"), (Q8_0, "q8_0")):
        w = np.zeros((m, k // 32 * WIRE[t]), np.uint8)
        r.ref_quantize_chunk(t, wf.ctypes.data_as(vp), w.ctypes.data_as(vp), C.c_int64(m), C.c_int64(k))
        g[f"w_{name}"] = w
        deq = np.zeros((m, k), np.float32)
        for i in range(m):
            r.ref_to_float(t, w[i].ctypes.data_as(vp), deq[i].ctypes.data_as(vp), C.c_int64(k))
        g[f"w_{name}_dequant"] = deq
        vd = np.zeros((m, x.shape[0]), np.float32)
        for i in range(m):
            for j in range(x.shape[0]):
                vd[i, j] = r.ref_vec_dot(t, C.c_int64(k), w[i].ctypes.data_as(vp), q8[j].ctypes.data_as(vp))
        g[f"vec_dot_{name}"] = vd
    g["w_f32"] = wf

    # --- MUL_MAT cases as test-backend-ops enumerates them: (type, m, n, k, bs0, bs1, nr0, nr1)
    cases = []
    for t in (Q4_0, Q8_0):
        for n in (1, 16):
            # same bs/nr pattern as the reference's list, batch 3 instead of 10 to keep the fixture small
            # (the full bs=10 list runs on the GPU through the reference's own test-backend-ops binary)
            cases.append((t, 16, n, 256, 1, 1, 1, 1))
            cases.append((t, 16, n, 256, 3, 1, 1, 1))
            cases.append((t, 16, n, 256, 3, 1, 2, 1))
            cases.append((t, 16, n, 256, 3, 2, 1, 1))
            cases.append((t, 16, n, 256, 3, 2, 2, 1))
            cases.append((t, 16, n, 256, 3, 2, 1, 2))
            cases.append((t, 16, n, 256, 3, 2, 2, 2))
        cases.append((t, 33, 5, 96, 1, 1, 1, 1))      # odd m, n, short k
        cases.append((t, 1, 1, 32, 1, 1, 1, 1))       # minimum
        cases.append((t, 130, 70, 128, 1, 1, 1, 1))   # crosses GEMM tile sizes
    g["mm_cases"] = np.array(cases, dtype=np.int64)
    for ci, (t, m, n, k, bs0, bs1, nr0, nr1) in enumerate(cases):
        a_f = rng.uniform(-1, 1, (bs1, bs0, m, k)).astype(np.float32)
        b_f = rng.uniform(-1, 1, (bs1 * nr1, bs0 * nr0, n, k)).astype(np.float32)
        a_q = np.zeros((bs1, bs0, m, k // 32 * WIRE[t]), np.uint8)
        r.ref_quantize_chunk(t, a_f.ctypes.data_as(vp), a_q.ctypes.data_as(vp), C.c_int64(bs1 * bs0 * m), C.c_int64(k))
        h = vp(r.ref_mm_create(t, C.c_int64(k), C.c_int64(m), C.c_int64(bs0), C.c_int64(bs1), C.c_int64(n),
                               C.c_int64(bs0 * nr0), C.c_int64(bs1 * nr1), 4))
        r.ref_mm_set_a(h, a_q.ctypes.data_as(vp))
        r.ref_mm_set_b(h, b_f.ctypes.data_as(vp))
        r.ref_mm_compute(h, 1)
        out = np.zeros((bs1 * nr1, bs0 * nr0, n, m), np.float32)
        r.ref_mm_get_out(h, out.ctypes.data_as(vp))
        r.ref_mm_free(h)
        g[f"mm{ci}_a"], g[f"mm{ci}_b"], g[f"mm{ci}_out"] = a_q, b_f.astype(np.float16), out
        # b is stored as fp16 to keep the fixture small; consumers must use b.astype(float32) (exact) as the input
        b32 = g[f"mm{ci}_b"].astype(np.float32)
        if not np.array_equal(b32, b_f):
            h = vp(r.ref_mm_create(t, C.c_int64(k), C.c_int64(m), C.c_int64(bs0), C.c_int64(bs1), C.c_int64(n),
                                   C.c_int64(bs0 * nr0), C.c_int64(bs1 * nr1), 4))
            r.ref_mm_set_a(h, a_q.ctypes.data_as(vp))
            r.ref_mm_set_b(h, b32.ctypes.data_as(vp))
            r.ref_mm_compute(h, 1)
            r.ref_mm_get_out(h, out.ctypes.data_as(vp))
            r.ref_mm_free(h)
            g[f"mm{ci}_out"] = out
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.1f} KiB, {len(g)} arrays)")


if __name__ == "__main__":
    main()
