mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_f16.py -x -q > gpurun_out/r02_gemm_tests.log 2>&1; tail -5 gpurun_out/r02_gemm_tests.log
timeout 200 python tools/ab_gemm.py > gpurun_out/r02_ab_gemm.log 2>&1; tail -20 gpurun_out/r02_ab_gemm.log
