mkdir -p gpurun_out
for shape in "4096 4096 512" "16384 4096 512" "4096 16384 512"; do echo "== $shape"; timeout 200 python tools/gemm_timeline.py q4_0 $shape 2>&1 | tail -9; done > gpurun_out/r02_gemm_timeline_gptj_shapes.log
cat gpurun_out/r02_gemm_timeline_gptj_shapes.log | cut -c1-160
