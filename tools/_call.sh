mkdir -p gpurun_out
L=ggml-imax_b200/lib/libggml_b200.so
cp $L /tmp/lib_ss.so
for rep in 1 2; do
  for v in ss ts; do
    if [ $v = ss ]; then cp /tmp/lib_ss.so $L; else cp tools/_build/libggml_b200_ts.so $L; fi
    timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v$rep.json 2> gpurun_out/ab_$v$rep.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v$rep.json').read().strip().splitlines()[-1])
t=d['targets']; print('$v$rep', 'C2 q4_0', t['c2_gemm_q4_0_us'], 'q8_0', t['c2_gemm_q8_0_us'], 'prefill512 ms', d['extra']['gptj6b_q4_0_prefill_512_tokens']['ms'], 'tok/s', d['value'])
PY
  done
done
cp /tmp/lib_ss.so $L
