#!/bin/bash
# Round-2q: full GPU suite with the row-per-thread GEMM epilogue (EF_DIRECT); A/B on the training / forward workloads.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02q_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02q_pytest.log | tail -8
b() { # tag workload [env...]
  local tag=$1; local wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-e2e --profile-out gpurun_out/r02q_${tag}_table.json > gpurun_out/r02q_bench_$tag.json 2> gpurun_out/r02q_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02q_bench_$tag.json').read().strip().splitlines()[-1])
    pk=d.get('per_kernel') or {}
    print('  $tag', d['value'], 'img/s', d['ms_per_step'], 'ms  gemm frac', d['roofline']['frac'], [(k[:24], round(v['ms'],2)) for k,v in list(pk.items())[:4]])
except Exception as e: print('  $tag no line', e)
PY
}
b fv4t_d0 fv4_train FVIT_GEMM_DIRECT=0
b fv4t_d1 fv4_train FVIT_GEMM_DIRECT=1
b fv0t_d0 fv0_train FVIT_GEMM_DIRECT=0
b fv0t_d1 fv0_train FVIT_GEMM_DIRECT=1
b fv4f_d0 fv4_fwd FVIT_GEMM_DIRECT=0
b fv4f_d1 fv4_fwd FVIT_GEMM_DIRECT=1
b fv0f_d0 fv0_fwd FVIT_GEMM_DIRECT=0
b fv0f_d1 fv0_fwd FVIT_GEMM_DIRECT=1
