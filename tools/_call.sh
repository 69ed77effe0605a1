mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_f16.py -x -q > gpurun_out/r02_gemm_tests.log 2>&1; tail -4 gpurun_out/r02_gemm_tests.log
for a in 4 0; do timeout 300 python tools/gemm_timeline.py q4_0 28672 8192 512 $a 2>&1 | grep -E "MMA thread|unit 0: MMAs|first MMA"; done
timeout 300 python tools/gemm_timeline.py q4_0 2>&1 | tail -9
