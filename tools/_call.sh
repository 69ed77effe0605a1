mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_gemm_f16.py tests/test_gpu_gptj_graph.py tests/test_gpu_gguf_load.py -q 2>&1 | tail -2
