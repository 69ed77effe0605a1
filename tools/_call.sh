mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gemm_f16.py -q -x > gpurun_out/r02_gemm_reduce_tests.log 2>&1; tail -2 gpurun_out/r02_gemm_reduce_tests.log
for shape in "4096 4096 512" "11008 4096 512"; do echo "== $shape"; timeout 200 python tools/gemm_timeline.py q4_0 $shape 2>&1 | tail -4; done
timeout 300 python tools/stress_gemm.py q4_0 4096 4096 512 60 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_reduce.json 2> gpurun_out/r02_bench_reduce.err; python - <<'PY'
import json
r=json.loads(open('gpurun_out/r02_bench_reduce.json').read().strip().splitlines()[-1])
x=r['extra']
print(r['value'], r['roofline']['frac'])
for k in ('c2_gemm_q4_0_m11008_k4096_n512','c2_gemm_q8_0_m11008_k4096_n512','gptj6b_q4_0_prefill_512_tokens'):
    print(k, {kk:vv for kk,vv in x[k].items() if kk in ('us_per_mul_mat','TFLOP/s','ms','prompt_tokens/s')})
PY
