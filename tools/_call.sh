mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02_full_gpu_suite.log 2>&1; tail -4 gpurun_out/r02_full_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_n1_now.json 2> gpurun_out/r02_bench_n1_now.err; cut -c1-300 gpurun_out/r02_bench_n1_now.json
for np in 8 32 128; do timeout 600 oracle/_ref/gptj-harness q4_0 28 4096 16 64 50400 2048 $np 2 $(nproc) > gpurun_out/r02_gptj_6b_p$np.json 2> gpurun_out/r02_gptj_6b.err; python - <<PY
import json
r=json.load(open('gpurun_out/r02_gptj_6b_p$np.json'))
for s in r['steps']: print({k:s[k] for k in ('n','logits_nmse_vs_cpu','ms_cpu','ms_b200','ms_b200_graph_plan','b200_launches')})
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 400 --csv --log-file gpurun_out/r02_gptj_decode_launches_final.csv oracle/_ref/gptj-harness q4_0 28 4096 16 64 50400 2048 8 1 16 1 0 > gpurun_out/ncu_gptj.log 2>&1; wc -l gpurun_out/r02_gptj_decode_launches_final.csv
