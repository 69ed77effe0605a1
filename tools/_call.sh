mkdir -p gpurun_out
cd oracle/_ref
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 9 ./test-backend-ops test -b B2000 > ../../gpurun_out/r02_sanitizer_backend_ops.log 2>&1; echo "test-backend-ops under memcheck rc $?"; tail -4 ../../gpurun_out/r02_sanitizer_backend_ops.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 9 ./gpt2-sched-harness q4_0 32 2 8 1 > ../../gpurun_out/r02_sanitizer_gpt2.log 2>&1; echo "gpt2 harness under memcheck rc $?"; tail -3 ../../gpurun_out/r02_sanitizer_gpt2.log | cut -c1-300
cd ../..
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 9 python -m pytest tests/test_gpu_ops.py tests/test_gpu_wire_formats.py -x -q > gpurun_out/r02_sanitizer_ops_tests.log 2>&1; echo "ops tests under memcheck rc $?"; tail -4 gpurun_out/r02_sanitizer_ops_tests.log | cut -c1-200
