#!/bin/bash
# Round-2k: full GPU suite, A/B of the GEMM tile cost models on fv4 / fv0 training and fv4 / ar0 forward.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02k_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|FAILED" gpurun_out/r02k_pytest.log | tail -8
b() { # tag workload [env...]
  local tag=$1; local wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-e2e --profile-out gpurun_out/r02k_${tag}_table.json > gpurun_out/r02k_bench_$tag.json 2> gpurun_out/r02k_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02k_bench_$tag.json').read().strip().splitlines()[-1])
    pk=d.get('per_kernel') or {}
    print('  $tag', d['value'], 'img/s', d['ms_per_step'], 'ms  gemm frac', d['roofline']['frac'], [(k, v['ms']) for k,v in list(pk.items())[:5]])
except Exception as e: print('  $tag no line', e)
PY
}
b fv4t_c0 fv4_train FVIT_GEMM_COST=0
b fv4t_c1 fv4_train FVIT_GEMM_COST=1
b fv0t_c0 fv0_train FVIT_GEMM_COST=0
b fv0t_c1 fv0_train FVIT_GEMM_COST=1
b fv0f_c0 fv0_fwd FVIT_GEMM_COST=0
b fv0f_c1 fv0_fwd FVIT_GEMM_COST=1
b ar0f_c0 ar0_fwd FVIT_GEMM_COST=0
b ar0f_c1 ar0_fwd FVIT_GEMM_COST=1
