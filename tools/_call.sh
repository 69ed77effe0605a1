mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gpt2_backend.py -x -q -s -k batched > gpurun_out/r02_gpt2_batched.log 2>&1; grep -E "gpt-2-|passed|failed" gpurun_out/r02_gpt2_batched.log | tail -4
python oracle/make_gpt2_model.py /tmp/m.bin q4_0 > /dev/null
P="the quick brown fox jumps over the lazy dog and keeps running through the forest until the night falls over the quiet hills of the north"
timeout 300 oracle/_ref/gpt-2-batched -m /tmp/m.bin -p "$P" -n 16 -s 7 --top_k 1 -b 256 -t 8 -np 4 -ngl 1 2>&1 | grep -E "time|n_decoded"
