#!/bin/bash
# Round-1d GPU validation + evidence: GPU tests, bench lines (fv4 / fv0 training) with per-launch tables, ncu launch
# list of one training step + optimizer step, ncu --set full of the optimizer kernels. Logs -> gpurun_out/r01d_*.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -p no:cacheprovider > gpurun_out/r01d_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; tail -25 gpurun_out/r01d_pytest.log | cut -c1-200
t2=$(date +%s)
timeout 200 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r01d_fv4_train_launch_table.json > gpurun_out/r01d_bench_fv4_train.json 2> gpurun_out/r01d_bench_fv4_train.err
echo "bench fv4 exit $? after $(( $(date +%s) - t2 ))s"; cut -c1-300 gpurun_out/r01d_bench_fv4_train.json; tail -3 gpurun_out/r01d_bench_fv4_train.err | cut -c1-300
t3=$(date +%s)
timeout 150 python bench.py --workload fv0_train --steps 10 --warmup 3 > gpurun_out/r01d_bench_fv0_train.json 2> gpurun_out/r01d_bench_fv0_train.err
echo "bench fv0 exit $? after $(( $(date +%s) - t3 ))s"; cut -c1-300 gpurun_out/r01d_bench_fv0_train.json; tail -3 gpurun_out/r01d_bench_fv0_train.err | cut -c1-300
python - <<'PY'
import json
for w in ("fv4", "fv0"):
    try:
        d = json.loads(open(f"gpurun_out/r01d_bench_{w}_train.json").read().strip().splitlines()[-1])
        print(w, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
        print(w, "optimizer_step:", json.dumps(d.get("optimizer_step"))[:900])
        pk = d.get("per_kernel") or {}
        print(w, "per_kernel:", [(k, v["ms"]) for k, v in list(pk.items())[:16]])
    except Exception as e:
        print(w, "no bench line:", e)
PY
t4=$(date +%s)
timeout 170 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r01d_launches_fv4_step.csv python scripts/gpu_train_step_profile.py \
    > gpurun_out/r01d_ncu_launches.log 2>&1
echo "ncu launch list exit $? after $(( $(date +%s) - t4 ))s"; tail -2 gpurun_out/r01d_ncu_launches.log | cut -c1-200; wc -l gpurun_out/r01d_launches_fv4_step.csv
t5=$(date +%s)
timeout 120 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:optim_ -c 8 \
    -o gpurun_out/r01d_optim python scripts/gpu_train_step_profile.py > gpurun_out/r01d_ncu_optim.log 2>&1
echo "ncu optim exit $? after $(( $(date +%s) - t5 ))s"; tail -2 gpurun_out/r01d_ncu_optim.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep 2>/dev/null
