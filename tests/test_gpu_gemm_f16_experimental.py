"""EXPERIMENTAL prefill path (B200_GEMM_F16=1, gemm_f16_kernel in b200_gemm_tc.cu): operands dequantized to fp16, fp32
accumulation over the whole k on the tensor cores -- what the reference's CUDA backend does for large batches
(ggml_cuda_op_mul_mat_cublas, src/ggml-cuda.cu:1208-1270).  The kernel was written at the end of round 1; only four small cases
(300,256,64 and 128,64,256, both types) have run on a GPU so far (they pass, profiles/r01_experimental_kernels_first_run.log), so these
tests stay skipped unless B200_TEST_EXPERIMENTAL=1 and the default `-m gpu` run only covers fully validated paths.  First thing to run in round 2:
    B200_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_gemm_f16_experimental.py -x -q
Bounds: NMSE <= 5e-4 against the oracle is the reference's bar (tests/test-backend-ops.cpp:921-923); fp16 rounding of both
operands (relative 2^-12 each) should leave it below 1e-6, and as close to the exact int8 kernel."""
import os

import numpy as np
import pytest

from conftest import Q4_0, Q8_0, MUL_MAT_NMSE_TOL, nmse

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200_TEST_EXPERIMENTAL") != "1", reason="unvalidated kernel: set B200_TEST_EXPERIMENTAL=1")]


@pytest.mark.parametrize("mode", ["1", "2"])      # 1: both operands materialised as fp16 (four small cases have run); 2: weights dequantized in the kernel (never run)
@pytest.mark.parametrize("qtype", [Q4_0, Q8_0])
@pytest.mark.parametrize("m,k,n", [(128, 64, 256), (128, 256, 256), (300, 256, 64), (1000, 4096, 512), (257, 96, 9), (64, 32, 300),
                                   (11008, 4096, 512)])
def test_f16_gemm_vs_oracle_and_exact_kernel(gpu_ctx, qmm, oracle, monkeypatch, qtype, m, k, n, mode):
    rng = np.random.default_rng(m * 31 + k + n)
    big = m * k > (1 << 24)
    wire = qmm.random_wire_weights(qtype, k, m, seed=m + k) if big else oracle.quantize_weights(qtype, rng.uniform(-1, 1, (m, k)).astype(np.float32))
    x = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    t = qmm.QTensor(gpu_ctx, qtype, k, m)
    try:
        t.set(wire)
        monkeypatch.delenv("B200_GEMM_F16", raising=False)
        exact = gpu_ctx.mul_mat(t, x, flags=qmm.MM_FORCE_GEMM)
        monkeypatch.setenv("B200_GEMM_F16", mode)
        got = gpu_ctx.mul_mat(t, x, flags=qmm.MM_FORCE_GEMM)
        assert got.shape == exact.shape and np.isfinite(got).all()
        assert nmse(got, exact) <= 1e-6, f"fp16 path vs exact int8 path: nmse {nmse(got, exact)}"
        rows = np.arange(m) if not big else np.unique(rng.integers(0, m, 64))
        ref = oracle.mul_mat(qtype, np.ascontiguousarray(wire[rows]), k, len(rows), 1, 1, x[None, None])[0, 0]
        assert nmse(got[:, rows], ref) <= MUL_MAT_NMSE_TOL
        assert nmse(got[:, rows], ref) <= 1e-6
    finally:
        t.free()
