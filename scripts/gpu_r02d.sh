#!/bin/bash
# Round-2d: A/B of the CTA-pair GEMM selection, fused HAT kernel and CUDA graphs on the bench workloads; new API tests.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
b() { # tag workload [env...]
  local tag=$1; local wl=$2; shift 2
  local t0=$(date +%s)
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-e2e --profile-out gpurun_out/r02d_${tag}_table.json > gpurun_out/r02d_bench_$tag.json 2> gpurun_out/r02d_bench_$tag.err
  echo "bench $tag exit $? after $(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/r02d_bench_$tag.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02d_bench_$tag.json').read().strip().splitlines()[-1])
    pk=d.get('per_kernel') or {}
    print('  $tag', d['value'], 'img/s', d['ms_per_step'], 'ms  gemm frac', d['roofline']['frac'], 'launches/step', d['launches_per_step'], [(k, v['ms']) for k,v in list(pk.items())[:6]])
except Exception as e: print('  $tag no line', e)
PY
}
b cg1 fv4_train FVIT_GEMM_CG=1
b auto08 fv4_train FVIT_GEMM_CG2_GAIN=0.8
b auto07 fv4_train FVIT_GEMM_CG2_GAIN=0.7
b auto09 fv4_train FVIT_GEMM_CG2_GAIN=0.9
b nograph fv4_train FVIT_CUDA_GRAPH=0
b fv0f_fused fv0_fwd FVIT_FUSED_HAT=1
b fv0f_unfused fv0_fwd FVIT_FUSED_HAT=0
b fv4f_fused fv4_fwd FVIT_FUSED_HAT=1
b fv4f_unfused fv4_fwd FVIT_FUSED_HAT=0
b ar0f ar0_fwd FVIT_FUSED_HAT=1
b fv0t fv0_train FVIT_FUSED_HAT=1
timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_ops_gpu.py tests/test_pack_gpu.py tests/test_optim_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r02d_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02d_pytest.log | cut -c1-230 | tail -20
