mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backend_ops.py tests/test_gpu_gpt2_sched.py tests/test_gpu_gpt2_backend.py -q > gpurun_out/r02_rope_graph_tests.log 2>&1; tail -15 gpurun_out/r02_rope_graph_tests.log
cd oracle/_ref && ./test-backend-ops test -b B2000 > ../../gpurun_out/r02_test_backend_ops_all.log 2>&1; tail -3 ../../gpurun_out/r02_test_backend_ops_all.log
