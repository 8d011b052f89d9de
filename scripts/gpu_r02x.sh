#!/bin/bash
# Round-2x: full GPU suite of the last tree (single-tile key-loop attention for 65..128-token windows in training: tiny_ar68
# golden, kernel tests with S = 68 / 100 / 104 / 128; mixin refactor of the training plan), smoke, short default bench.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r02x_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02x_pytest.log | tail -24
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02x_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/r02x_smoke.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 4 --no-also --no-e2e > gpurun_out/r02x_bench_fv4t.json 2> gpurun_out/r02x_bench_fv4t.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02x_bench_fv4t.json').read().strip().splitlines()[-1])
    print('fv4_train', d['value'], 'img/s', d['ms_per_step'], 'ms roofline', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], d['clocks'])
except Exception as e: print('no bench line', e)
PY
