mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gpt2_backend.py -x -q -s -k "simple or test_mul_mat" > gpurun_out/r02_ref_examples.log 2>&1; tail -5 gpurun_out/r02_ref_examples.log
oracle/_ref/test-mul-mat 2>&1 | tail -8
