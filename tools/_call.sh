mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gptj_graph.py -x -q > gpurun_out/r02_gptj_tests.log 2>&1; tail -12 gpurun_out/r02_gptj_tests.log
timeout 600 oracle/_ref/gptj-harness q4_0 28 4096 16 64 50400 2048 8 3 $(nproc) > gpurun_out/r02_gptj_6b.json 2> gpurun_out/r02_gptj_6b.err; cut -c1-1500 gpurun_out/r02_gptj_6b.json; tail -3 gpurun_out/r02_gptj_6b.err
