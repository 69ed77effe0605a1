mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 400 $TR bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -3 gpurun_out/r02_bench_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_n$N.json
timeout 400 $TR bench.py --gpus $N --steps 200 --warmup 5 --workload c5 > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err; tail -3 gpurun_out/r02_bench_c5_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_c5_n$N.json
