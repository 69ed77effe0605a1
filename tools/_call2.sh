mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin_graph.py -q > gpurun_out/r02_multi_tests.log 2>&1; tail -4 gpurun_out/r02_multi_tests.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -2 gpurun_out/r02_bench_n2.err | cut -c1-300
python - <<PY
import json
for f in ('gpurun_out/r02_bench_n2.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['how']['path'], {k:v for k,v in d['checks'].items() if 'note' not in k and 'kind' not in k and 'nmse_same' not in k})
    except Exception as e: print(f, 'ERR', e)
PY
