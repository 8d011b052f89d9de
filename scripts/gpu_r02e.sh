#!/bin/bash
# Round-2e: 8-wide BN / affine kernels, backbone + deploy tests, full GPU suite, fv4 bench.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider --durations=6 > gpurun_out/r02e_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error|^[0-9.]+s " gpurun_out/r02e_pytest.log | cut -c1-230 | tail -25
timeout 300 python bench.py --workload fv4_train --steps 10 --warmup 3 --no-also --no-e2e --profile-out gpurun_out/r02e_fv4_train_table.json > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
echo "bench exit $?"; tail -2 gpurun_out/r02e_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
pk=d.get('per_kernel') or {}
print(d['value'], 'img/s', d['ms_per_step'], 'ms  gemm frac', d['roofline']['frac'], [(k, v['ms']) for k,v in list(pk.items())[:12]])
PY
