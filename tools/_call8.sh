mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --no-extras > gpurun_out/r02_bench_final_n$N.json 2> gpurun_out/r02_bench_final_n$N.err; tail -2 gpurun_out/r02_bench_final_n$N.err | cut -c1-300
python - <<PY
import json
for f in ('gpurun_out/r02_bench_final_n$N.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], {k:v for k,v in d['checks'].items() if 'note' not in k and 'kind' not in k})
    except Exception as e: print(f, 'ERR', e)
PY
