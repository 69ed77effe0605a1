mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -q -x > gpurun_out/r02_gemv_ncols_tests.log 2>&1; tail -4 gpurun_out/r02_gemv_ncols_tests.log
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:gemv_stream --csv --log-file gpurun_out/r02_gemv_ncols_after.csv python tools/_gemv8_once.py > gpurun_out/ncu_gemv8.log 2>&1
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/r02_gemv_ncols_after.csv') if l.startswith('"')]
for x in csv.DictReader(lines):
    print(re.sub(r'\(.*','',x['Kernel Name'])[-40:], x['Metric Name'], x['Metric Value'])
PY
