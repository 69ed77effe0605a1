mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; tail -4 gpurun_out/r02_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['targets']['c2_gemm_q4_0_us'], d['targets']['c2_gemm_q8_0_us'], d['extra']['gptj6b_q4_0_prefill_512_tokens']['ms'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 600 gpurun_out/r02_bench_ref.json
timeout 600 oracle/_ref/gpt2-sched-harness q4_0 128 4 8 1 > gpurun_out/plain_gpt2.json 2> gpurun_out/plain_gpt2.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 470 --launch-count 360 --csv --log-file gpurun_out/r02_gpt2_decode_launches.csv oracle/_ref/gpt2-sched-harness q4_0 128 4 8 1 > gpurun_out/ncu_gpt2.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_gpt2_decode_launches.csv')) if len(r)>14 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0].replace('void <unnamed>::','')
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=float(r[14])/1e3
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:60s} {v[0]:5d} {v[1]:10.1f} us  {v[1]/v[0]:7.2f} us/launch {v[1]/tot:6.3f}")
print(len(rows), tot)
PY
