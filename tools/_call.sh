mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gptj_graph.py tests/test_gpu_gpt2_sched.py -q 2>&1 | tail -2
timeout 600 oracle/_ref/gptj-harness q4_0 28 4096 16 64 50400 2048 8 3 $(nproc) > gpurun_out/r02_gptj_6b_p8.json 2> gpurun_out/r02_gptj_6b.err; python - <<PY
import json
r=json.load(open('gpurun_out/r02_gptj_6b_p8.json'))
print([(s['n'], s['ms_b200'], s['ms_b200_graph_plan'], s['graph_plan_equals_node_by_node']) for s in r['steps']], r['ok'])
PY
timeout 300 oracle/_ref/gpt2-sched-harness q4_0 128 3 8 1 0 > gpurun_out/r02_gpt2_plan.json; python - <<'PY'
import json
r=json.load(open('gpurun_out/r02_gpt2_plan.json'))
print([(s['n'], s['ms_b200_whole_graph'], s['ms_b200_graph_plan'], s['graph_plan_equals_node_by_node']) for s in r['steps']], r['ok'])
PY
