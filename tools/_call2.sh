mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02_multi_tests.log 2>&1; tail -15 gpurun_out/r02_multi_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -4 gpurun_out/r02_bench_n2.err; cut -c1-1800 gpurun_out/r02_bench_n2.json
timeout 400 $TR bench.py --gpus 2 --steps 200 --warmup 5 --workload c5 > gpurun_out/r02_bench_c5_n2.json 2> gpurun_out/r02_bench_c5_n2.err; tail -4 gpurun_out/r02_bench_c5_n2.err; cut -c1-1600 gpurun_out/r02_bench_c5_n2.json
timeout 300 $TR tools/prefill_rowsplit.py > gpurun_out/r02_prefill_rowsplit_n2.json 2> gpurun_out/r02_prefill_rowsplit_n2.err; tail -3 gpurun_out/r02_prefill_rowsplit_n2.err; cat gpurun_out/r02_prefill_rowsplit_n2.json
timeout 200 python tools/prefill_rowsplit.py > gpurun_out/r02_prefill_rowsplit_n1.json 2> gpurun_out/r02_prefill_rowsplit_n1.err; cat gpurun_out/r02_prefill_rowsplit_n1.json
