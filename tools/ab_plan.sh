run() { label="$1"; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab.json 2> gpurun_out/ab.err; python -c "
import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$label', d['value'], d['e2e']['value'], d['config']['plan_vs_launch_per_node_bitwise'])"; }
run "new-order no-act 8slots" B200_PLAN_NO_ACT_WARPS=1
run "new-order no-act 10slots" B200_PLAN_SLOTS=10
run "old-order no-act 8slots" B200_BENCH_ORDER=qkv_first
run "old-order no-act 10slots" B200_BENCH_ORDER=qkv_first B200_PLAN_SLOTS=10
run "old-order launches" B200_BENCH_ORDER=qkv_first B200_X=1
