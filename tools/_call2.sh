mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "k_split or row_split" > gpurun_out/r02_multi_tests.log 2>&1; tail -3 gpurun_out/r02_multi_tests.log | cut -c1-600
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_ks_n1.json 2> gpurun_out/r02_bench_ks_n1.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_ks_n1.json').read().strip().splitlines()[-1]); print('N=1', d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
bash tools/_call8.sh 2
