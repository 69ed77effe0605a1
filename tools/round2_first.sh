#!/bin/bash
# First GPU call of round 2 (one B200, ~3 minutes):  gpurun --timeout 400 -- 'bash tools/round2_first.sh'
# Runs the gated tests of the experimental kernels, then times them against the shipped ones.
# Everything lands in gpurun_out/r02_first_*.log; nothing here changes a default.
mkdir -p gpurun_out
export B200_TEST_EXPERIMENTAL=1
timeout 150 python -m pytest tests/test_gpu_plan.py -q -k "published_planes or without_k_split" > gpurun_out/r02_first_plan_tests.log 2>&1
tail -3 gpurun_out/r02_first_plan_tests.log
timeout 150 python -m pytest tests/test_gpu_gemm_f16_experimental.py -q > gpurun_out/r02_first_gemm_tests.log 2>&1
tail -3 gpurun_out/r02_first_gemm_tests.log
unset B200_TEST_EXPERIMENTAL
timeout 100 python tools/ab_ring.py pubq > gpurun_out/r02_first_ab_plan.log 2>&1
grep -E "us/token|bitwise|Error|error" gpurun_out/r02_first_ab_plan.log | cut -c1-160
timeout 120 python tools/ab_gemm.py > gpurun_out/r02_first_ab_gemm.log 2>&1
cat gpurun_out/r02_first_ab_gemm.log | tail -8
