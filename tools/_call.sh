mkdir -p gpurun_out
(cd oracle/_ref && timeout 600 ./test-backend-ops test -b B2000 -o MUL_MAT_ID > ../../gpurun_out/r02_backend_ops_mmid.log 2>&1; echo "rc $?"); sed 's/\x1b\[[0-9;]*m//g' gpurun_out/r02_backend_ops_mmid.log | grep -v "not supported" | tail -12; sed 's/\x1b\[[0-9;]*m//g' gpurun_out/r02_backend_ops_mmid.log | grep -c " OK$"
timeout 900 python -m pytest tests/test_gpu_backend_ops.py -x -q > gpurun_out/r02_spi_tests.log 2>&1; tail -4 gpurun_out/r02_spi_tests.log
(cd oracle/_ref && timeout 600 ./test-backend-ops test -b B2000 > ../../gpurun_out/r02_backend_ops_all.log 2>&1); sed 's/\x1b\[[0-9;]*m//g' gpurun_out/r02_backend_ops_all.log | grep -c " OK$"
