mkdir -p gpurun_out
timeout 400 python tools/ab_plan.py sweep > gpurun_out/r02_ab_plan3.log 2>&1; grep -E "us/token|bitwise|Error|error" gpurun_out/r02_ab_plan3.log | cut -c1-250
