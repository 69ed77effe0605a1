mkdir -p gpurun_out
timeout 600 oracle/_ref/gpt2-sched-harness q4_0 128 3 8 > gpurun_out/r02_gpt2_sched_q4_0.json 2> gpurun_out/r02_gpt2_sched_q4_0.err; cat gpurun_out/r02_gpt2_sched_q4_0.json | cut -c1-900; tail -3 gpurun_out/r02_gpt2_sched_q4_0.err
timeout 600 python -m pytest tests/test_gpu_gpt2_sched.py tests/test_gpu_dropin_graph.py tests/test_gpu_backend_ops.py tests/test_gpu_plan.py -x -q > gpurun_out/r02_spi_tests.log 2>&1; tail -8 gpurun_out/r02_spi_tests.log
