# A/B of the bench under the plan kernel's knobs (run on the GPU box: bash tools/ab_plan.sh)
run() { label="$1"; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab.json 2> gpurun_out/ab.err; python -c "
import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$label', d['value'], d['e2e']['value'], d['config']['plan_vs_launch_per_node_bitwise'])"; }
run "plan, fc_in-first order (default)" X=1
run "plan, q-k-v-first order" B200_BENCH_ORDER=qkv_first
run "plan, 6 ring slots" B200_PLAN_SLOTS=6
run "plan, L2 prefetch 2 ops ahead" B200_PLAN_L2_AHEAD=2
