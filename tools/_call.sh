mkdir -p gpurun_out
for s in "q4_0 4096 4096 512" "q4_0 4096 16384 128" "q8_0 11008 4096 512" "q4_0 2304 768 128"; do timeout 200 python tools/stress_gemm.py $s 150 2>&1 | tail -4; done > gpurun_out/r02_stress_gemm.log 2>&1
cat gpurun_out/r02_stress_gemm.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -5 gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -5 gpurun_out/r02_bench_n1.err; cut -c1-2500 gpurun_out/r02_bench_n1.json
timeout 300 python bench.py --workload c5 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err; tail -3 gpurun_out/r02_bench_c5_n1.err; cut -c1-1500 gpurun_out/r02_bench_c5_n1.json
