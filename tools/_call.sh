mkdir -p gpurun_out
for np in 8 32 128; do timeout 600 oracle/_ref/gptj-harness q4_0 28 4096 16 64 50400 2048 $np 2 $(nproc) > gpurun_out/r02_gptj_6b_p$np.json 2> gpurun_out/r02_gptj_6b.err; python - <<PY
import json
r=json.load(open('gpurun_out/r02_gptj_6b_p$np.json'))
for s in r['steps']: print({k:s[k] for k in ('n','logits_nmse_vs_cpu','ms_cpu','ms_b200','ms_b200_first_call','ms_b200_graph_plan','b200_launches')})
print(r['ok'])
PY
done
