#!/bin/bash
# Round-2f: A/B of 12 epilogue warps (variant library), bn_bwd v8/v4, attention-backward TMEM split.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
b() { # tag workload [env...]
  local tag=$1; local wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-e2e --profile-out gpurun_out/r02f_${tag}_table.json > gpurun_out/r02f_bench_$tag.json 2> gpurun_out/r02f_bench_$tag.err
  echo "bench $tag exit $?"; tail -1 gpurun_out/r02f_bench_$tag.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02f_bench_$tag.json').read().strip().splitlines()[-1])
    pk=d.get('per_kernel') or {}
    print('  $tag', d['value'], 'img/s', d['ms_per_step'], 'ms  gemm frac', d['roofline']['frac'], [(k, v['ms']) for k,v in list(pk.items())[:9]])
except Exception as e: print('  $tag no line', e)
PY
}
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_train_ops_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02f_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|FAILED" gpurun_out/r02f_pytest.log | tail -8
FVIT_LIB=$PWD/fastervit_b200/libfvit_sm100_e12.so timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_train_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02f_pytest_e12.log 2>&1
echo "pytest e12 exit $?"; grep -E "passed|failed|FAILED" gpurun_out/r02f_pytest_e12.log | tail -8
b base fv4_train A=1
b bnv4 fv4_train FVIT_BN_V8=0
b e12 fv4_train FVIT_LIB=$PWD/fastervit_b200/libfvit_sm100_e12.so
b base2 fv4_train A=1
b e12b fv4_train FVIT_LIB=$PWD/fastervit_b200/libfvit_sm100_e12.so
b fv0t_e12 fv0_train FVIT_LIB=$PWD/fastervit_b200/libfvit_sm100_e12.so
b fv0t fv0_train A=1
