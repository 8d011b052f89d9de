#!/bin/bash
# Round-1e final GPU confirmation: GPU tests, smoke, bench lines of all five workloads. Logs -> gpurun_out/r01e_*.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 240 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -p no:cacheprovider > gpurun_out/r01e_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; tail -22 gpurun_out/r01e_pytest.log | cut -c1-200
timeout 100 python __graft_entry__.py smoke > gpurun_out/r01e_smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/r01e_smoke.log | cut -c1-300
for w in fv4_train fv0_train fv0_fwd fv4_fwd ar0_fwd; do
  t1=$(date +%s)
  extra=""
  [ "$w" = "fv4_train" ] && extra="--profile-out gpurun_out/r01e_fv4_train_launch_table.json"
  timeout 150 python bench.py --workload $w --steps 10 --warmup 3 $extra > gpurun_out/r01e_bench_$w.json 2> gpurun_out/r01e_bench_$w.err
  echo "bench $w exit $? after $(( $(date +%s) - t1 ))s"; tail -2 gpurun_out/r01e_bench_$w.err | cut -c1-200
done
python - <<'PY'
import json
for w in ("fv4_train", "fv0_train", "fv0_fwd", "fv4_fwd", "ar0_fwd"):
    try:
        d = json.loads(open(f"gpurun_out/r01e_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", d["value"], "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1), "gemm frac", d["roofline"]["frac"],
              "cpu", round(d["cpu_baseline"]["value"], 2), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        if d.get("optimizer_step"):
            print("   optimizer:", json.dumps({k: v for k, v in d["optimizer_step"].items() if k in ("fused_ema", "separate_ema")})[:600])
        pk = d.get("per_kernel") or {}
        print("   per_kernel:", [(k, v["ms"]) for k, v in list(pk.items())[:12]])
    except Exception as e:
        print(w, "no bench line:", e)
PY
