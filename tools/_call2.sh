mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k split_buffer > gpurun_out/r02_multi_tests.log 2>&1; tail -12 gpurun_out/r02_multi_tests.log | cut -c1-400
