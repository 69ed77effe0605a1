mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; tail -6 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
