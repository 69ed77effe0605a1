mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gptj_graph.py tests/test_gpu_gpt2_sched.py tests/test_gpu_gpt2_backend.py -x -q > gpurun_out/r02_gptj_tests.log 2>&1; tail -12 gpurun_out/r02_gptj_tests.log
timeout 600 oracle/_ref/gptj-harness q4_0 28 4096 16 64 50400 2048 8 3 $(nproc) > gpurun_out/r02_gptj_6b.json 2> gpurun_out/r02_gptj_6b.err; cut -c1-2500 gpurun_out/r02_gptj_6b.json; tail -3 gpurun_out/r02_gptj_6b.err
timeout 300 oracle/_ref/gpt2-sched-harness q4_0 128 3 8 1 0 > gpurun_out/r02_gpt2_plan.json; python - <<'PY'
import json
r=json.load(open('gpurun_out/r02_gpt2_plan.json'))
for s in r['steps']: print({k:s[k] for k in ('n','b200_whole_graph_logits_nmse_vs_cpu','ms_b200_whole_graph','ms_b200_graph_plan','b200_launches','graph_plan_equals_node_by_node')})
PY
