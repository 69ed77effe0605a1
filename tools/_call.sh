mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gpt2_sched.py tests/test_gpu_gpt2_backend.py -x -q > gpurun_out/r02_spi_tests.log 2>&1; tail -5 gpurun_out/r02_spi_tests.log
timeout 300 oracle/_ref/gpt2-sched-harness q4_0 128 3 8 1 1 | cut -c1-600
