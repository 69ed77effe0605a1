mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gemm_f16.py tests/test_gpu_parity.py -q > gpurun_out/r02_smalln_tests.log 2>&1; tail -5 gpurun_out/r02_smalln_tests.log
for shape in "4096 4096 16" "11008 4096 512"; do echo "== $shape"; timeout 200 python tools/gemm_timeline.py q4_0 $shape 2>&1 | tail -9; done > gpurun_out/r02_gemm_timeline_smalln_after.log
cat gpurun_out/r02_gemm_timeline_smalln_after.log | cut -c1-200
timeout 600 python tools/time_gemm.py 2>&1 | tail -3
