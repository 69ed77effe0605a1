mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_f16.py -x -q > gpurun_out/r02_gemm_tests.log 2>&1; tail -4 gpurun_out/r02_gemm_tests.log
for a in 0 2; do timeout 300 python tools/gemm_timeline.py q4_0 28672 8192 512 $a 2>&1 | grep -E "launch|first MMA|unit 0: MMAs"; done
timeout 300 python tools/gemm_timeline.py q4_0 2>&1 | tail -9
timeout 600 python tools/ab_gemm.py > gpurun_out/r02_ab_gemm_ts2.log 2>&1; cat gpurun_out/r02_ab_gemm_ts2.log | tail -20
