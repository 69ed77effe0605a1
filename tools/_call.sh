mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_f16.py -x -q > gpurun_out/r02_gemm_tests.log 2>&1; tail -4 gpurun_out/r02_gemm_tests.log
timeout 300 python tools/gemm_timeline.py q4_0 2>&1 | tail -9
timeout 300 python tools/gemm_timeline.py q4_0 28672 8192 512 2>&1 | grep -E "first MMA|unit 0: MMAs|end"
timeout 600 python tools/ab_gemm.py > gpurun_out/r02_ab_gemm_k128.log 2>&1; cat gpurun_out/r02_ab_gemm_k128.log | tail -20
