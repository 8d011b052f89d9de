#!/bin/bash
# Round-1c GPU validation: full GPU test suite, smoke(), the default bench line. Logs -> gpurun_out/r01c_*.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r01c_gpus.txt 2>&1
t0=$(date +%s)
timeout 420 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 -p no:cacheprovider > gpurun_out/r01c_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; tail -30 gpurun_out/r01c_pytest.log | cut -c1-220
t1=$(date +%s)
timeout 120 python __graft_entry__.py smoke > gpurun_out/r01c_smoke.log 2>&1
echo "smoke exit $? after $(( $(date +%s) - t1 ))s"; tail -6 gpurun_out/r01c_smoke.log | cut -c1-300
t2=$(date +%s)
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/r01c_bench_fv4_train.json 2> gpurun_out/r01c_bench_fv4_train.err
echo "bench exit $? after $(( $(date +%s) - t2 ))s"; cut -c1-900 gpurun_out/r01c_bench_fv4_train.json; tail -5 gpurun_out/r01c_bench_fv4_train.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r01c_bench_fv4_train.json").read().strip().splitlines()[-1])
    print("optimizer_step:", json.dumps(d.get("optimizer_step"))[:1200])
    pk = d.get("per_kernel") or {}
    print("per_kernel top:", [(k, v["ms"]) for k, v in list(pk.items())[:14]])
    print("roofline:", d.get("roofline"))
except Exception as e:
    print("no bench line:", e)
PY
