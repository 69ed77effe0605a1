mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02_full_gpu_suite.log 2>&1; tail -3 gpurun_out/r02_full_gpu_suite.log
