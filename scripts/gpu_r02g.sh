#!/bin/bash
# Round-2g: 12 epilogue warps default, attention backward with register dbias; tests + benches.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02g_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|FAILED" gpurun_out/r02g_pytest.log | tail -8
for wl in fv4_train fv0_train; do
timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-e2e --profile-out gpurun_out/r02g_${wl}_table.json > gpurun_out/r02g_bench_$wl.json 2> gpurun_out/r02g_bench_$wl.err
echo "bench $wl exit $?"; tail -1 gpurun_out/r02g_bench_$wl.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench_$wl.json').read().strip().splitlines()[-1])
pk=d.get('per_kernel') or {}
print('  $wl', d['value'], 'img/s', d['ms_per_step'], 'ms  gemm frac', d['roofline']['frac'], [(k, v['ms']) for k,v in list(pk.items())[:14]])
PY
done
