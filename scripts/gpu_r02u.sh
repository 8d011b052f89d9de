#!/bin/bash
# Round-2u: full GPU suite, smoke, the default bench line (with e2e, cpu_baseline, also_measured incl. batch-8 latency and the 21k-384 training step), reference arm -- weight-gradient side branch on, long-window attention backward for every S > 128.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --durations=5 > gpurun_out/r02u_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02u_pytest.log | tail -8
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02u_smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/r02u_smoke.log | cut -c1-300
t1=$(date +%s)
timeout 500 python bench.py --profile-out gpurun_out/r02u_fv4_train_launch_table.json > gpurun_out/r02u_bench_default.json 2> gpurun_out/r02u_bench_default.err
echo "default bench exit $? after $(( $(date +%s) - t1 ))s"; tail -2 gpurun_out/r02u_bench_default.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
print('also', json.dumps(d['also_measured'])[:1500])
PY
t2=$(date +%s)
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02u_bench_reference.json 2> gpurun_out/r02u_bench_reference.err
echo "reference arm exit $? after $(( $(date +%s) - t2 ))s"; cut -c1-700 gpurun_out/r02u_bench_reference.json
