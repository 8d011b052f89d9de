#!/bin/bash
# Round-2a: new parity tests (fv4 / S=85 / S=148 backward, per-module activations, training-op units), baseline bench.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 --durations=8 -p no:cacheprovider -s > gpurun_out/r02a_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error|worst activation|median rel" gpurun_out/r02a_pytest.log | cut -c1-260 | tail -60
for w in fv4_train; do
  t1=$(date +%s)
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --profile-out gpurun_out/r02a_${w}_launch_table.json > gpurun_out/r02a_bench_$w.json 2> gpurun_out/r02a_bench_$w.err
  echo "bench $w exit $? after $(( $(date +%s) - t1 ))s"; tail -2 gpurun_out/r02a_bench_$w.err | cut -c1-200
  python -c "
import json
d=json.loads(open('gpurun_out/r02a_bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])
pk=d.get('per_kernel') or {}
print([(k, v['ms']) for k,v in list(pk.items())[:14]])
"
done
