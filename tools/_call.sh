mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02_full_gpu_suite.log 2>&1; tail -5 gpurun_out/r02_full_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1_now.json 2> gpurun_out/r02_bench_n1_now.err; cut -c1-400 gpurun_out/r02_bench_n1_now.json
python bench.py --impl reference > gpurun_out/r02_bench_ref_now.json 2> gpurun_out/r02_bench_ref_now.err; cut -c1-500 gpurun_out/r02_bench_ref_now.json
