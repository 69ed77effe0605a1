mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/r02_ops_tests.log 2>&1; tail -15 gpurun_out/r02_ops_tests.log
for f in 1 0; do timeout 600 oracle/_ref/gpt2-sched-harness q4_0 128 8 8 $f > gpurun_out/r02_gpt2_q4_0_fuse$f.json 2> gpurun_out/r02_gpt2_q4_0.err; python - <<PY
import json
r=json.loads(open('gpurun_out/r02_gpt2_q4_0_fuse$f.json').read().strip().splitlines()[-1])
print('fuse=$f', 'ok', r.get('ok'), 'fused', r.get('b200_fused_nodes_total'), [(s['n'], s['ms_b200_whole_graph'], s['b200_launches'], s['graph_nodes'], '%.2e'%s['b200_whole_graph_logits_nmse_vs_cpu']) for s in r['steps']])
PY
done
tail -5 gpurun_out/r02_gpt2_q4_0.err
timeout 900 python -m pytest tests/test_gpu_gpt2_backend.py tests/test_gpu_gpt2_sched.py tests/test_gpu_backend_ops.py tests/test_gpu_dropin_graph.py -x -q -s > gpurun_out/r02_gpt2_backend.log 2>&1; grep -E "gpt-2-backend|passed|failed|Error" gpurun_out/r02_gpt2_backend.log | tail -12
