#!/bin/bash
# Round-2o: full GPU suite with (a) leaf launches as side branches of the launch graphs and (b) zero-filled K steps
# not issued in the GEMM main loop; A/B of both on the fv4 / fv0 training steps; A-operand row-stride micro-benchmark.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02o_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02o_pytest.log | tail -8
b() { # tag workload [env...]
  local tag=$1; local wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-e2e > gpurun_out/r02o_bench_$tag.json 2> gpurun_out/r02o_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02o_bench_$tag.json').read().strip().splitlines()[-1])
    pk=d.get('per_kernel') or {}
    print('  $tag', d['value'], 'img/s', d['ms_per_step'], 'ms  gemm frac', d['roofline']['frac'], [(k[:24], round(v['ms'],2)) for k,v in list(pk.items())[:6]])
except Exception as e: print('  $tag no line', e)
PY
}
b fv4t_base fv4_train FVIT_SIDE_BRANCHES=0 FVIT_GEMM_KSKIP=0
b fv4t_kskip fv4_train FVIT_SIDE_BRANCHES=0 FVIT_GEMM_KSKIP=1
b fv4t_prep fv4_train FVIT_SIDE_BRANCHES=2 FVIT_GEMM_KSKIP=1
b fv4t_leaf fv4_train FVIT_SIDE_BRANCHES=3 FVIT_GEMM_KSKIP=1
b fv4t_all fv4_train FVIT_SIDE_BRANCHES=1 FVIT_GEMM_KSKIP=1
b fv0t_base fv0_train FVIT_SIDE_BRANCHES=0 FVIT_GEMM_KSKIP=0
b fv0t_all fv0_train FVIT_SIDE_BRANCHES=1 FVIT_GEMM_KSKIP=1
timeout 200 python scripts/gpu_gemm_stride_micro.py 2>&1 | tee gpurun_out/r02o_stride_micro.txt
