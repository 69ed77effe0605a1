mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; tail -6 gpurun_out/r02_gpu_tests.log
timeout 600 python tools/stress_gemm.py > gpurun_out/r02_stress_gemm.log 2>&1; tail -6 gpurun_out/r02_stress_gemm.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['targets'], d['checks'].get('logits_vs_oracle_nmse'), d['clocks'])
print(d['extra'].get('gptj6b_q4_0_prefill_512_tokens'))
PY
