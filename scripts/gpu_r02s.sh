#!/bin/bash
# Round-2s: weight-gradient GEMMs as side-branch launches (FVIT_WGRAD_SIDE=1): gradient parity on every training golden,
# then A/B on the two training workloads.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t0=$(date +%s)
FVIT_WGRAD_SIDE=1 timeout 600 python -m pytest tests/test_train_gpu.py tests/test_optim_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02s_pytest.log 2>&1
echo "pytest (FVIT_WGRAD_SIDE=1) exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02s_pytest.log | tail -8
b() { # tag workload [env...]
  local tag=$1; local wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 12 --warmup 4 --no-also --no-e2e > gpurun_out/r02s_bench_$tag.json 2> gpurun_out/r02s_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02s_bench_$tag.json').read().strip().splitlines()[-1])
    print('  $tag', d['value'], 'img/s', d['ms_per_step'], 'ms  clocks', d['clocks'].get('sm_mhz'))
except Exception as e: print('  $tag no line', e)
PY
}
b fv4t_s0 fv4_train FVIT_WGRAD_SIDE=0
b fv4t_s1 fv4_train FVIT_WGRAD_SIDE=1
b fv4t_s0b fv4_train FVIT_WGRAD_SIDE=0
b fv4t_s1b fv4_train FVIT_WGRAD_SIDE=1
b fv0t_s0 fv0_train FVIT_WGRAD_SIDE=0
b fv0t_s1 fv0_train FVIT_WGRAD_SIDE=1
# long-window attention backward: cost split and one ncu --set full capture
timeout 200 python scripts/gpu_attn_long_micro.py > gpurun_out/r02s_attn_long_micro.txt 2>&1; cat gpurun_out/r02s_attn_long_micro.txt | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_loop_bwd_long -s 2 -c 1 \
    -o gpurun_out/r02s_attn_long python scripts/gpu_attn_long_micro.py > gpurun_out/r02s_ncu_attn_long.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r02s_ncu_attn_long.log | cut -c1-160
