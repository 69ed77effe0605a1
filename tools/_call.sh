mkdir -p gpurun_out
timeout 600 oracle/_ref/gptj-harness q8_0 28 4096 16 64 50400 2048 8 3 $(nproc) > gpurun_out/r02_gptj_6b_q8_0.json 2> gpurun_out/r02_gptj_6b.err; python - <<PY
import json
r=json.load(open('gpurun_out/r02_gptj_6b_q8_0.json'))
print(r['mul_mat_weight_bytes_per_token'], [(s['n'], s['logits_nmse_vs_cpu'], s['ms_cpu'], s['ms_b200'], s['ms_b200_graph_plan'], s['b200_launches']) for s in r['steps']], r['ok'])
PY
tail -2 gpurun_out/r02_gptj_6b.err
