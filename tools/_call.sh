mkdir -p gpurun_out
(cd oracle/_ref && timeout 900 ./test-backend-ops perf -b B2000 > ../../gpurun_out/r02_backend_ops_perf.log 2>&1; echo "rc $?")
sed 's/\x1b\[[0-9;]*m//g' gpurun_out/r02_backend_ops_perf.log | grep -v "not supported" | grep -B1 "runs" | grep -v "^--" | paste - - | awk '{print}' | sort -t'-' -k4 | tail -25 | cut -c1-220
sed 's/\x1b\[[0-9;]*m//g' gpurun_out/r02_backend_ops_perf.log | grep -c "GB/s"
