mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gguf_load.py -x -q > gpurun_out/r02_gguf_tests.log 2>&1; tail -12 gpurun_out/r02_gguf_tests.log
timeout 600 oracle/_ref/gguf-load-harness 4096 8 /tmp/gptj8.gguf > gpurun_out/r02_gguf_load.json 2> gpurun_out/r02_gguf_load.err; cat gpurun_out/r02_gguf_load.json; tail -3 gpurun_out/r02_gguf_load.err
