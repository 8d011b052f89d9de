#!/bin/bash
# Round-2n: full GPU suite after the loop-attention producer-order fix; ar0 / fv4 forward lines.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --durations=5 > gpurun_out/r02n_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02n_pytest.log | tail -8
for wl in ar0_fwd; do
timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-also > gpurun_out/r02n_bench_$wl.json 2> gpurun_out/r02n_bench_$wl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02n_bench_$wl.json').read().strip().splitlines()[-1])
pk=d.get('per_kernel') or {}
print('  $wl', d['value'], 'img/s', d['ms_per_step'], 'ms e2e', d['e2e']['value'], [(k, v['ms']) for k,v in list(pk.items())[:6]])
PY
done
