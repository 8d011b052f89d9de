"""profiles/<tag>_traffic.json from an ncu launch list with DRAM bytes (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch): DRAM bytes per fvit_gemm launch (what bench.py reports as roofline.traffic), the GEMM
share of device time under ncu, and the whole step's DRAM traffic.

    python scripts/traffic_from_launches.py r02i gpurun_out/r02i_launches_fv4_step.csv fv4_train
"""
import collections
import csv
import io
import json
import sys
from pathlib import Path

tag, path, workload = sys.argv[1], Path(sys.argv[2]), sys.argv[3]
lines = [l for l in path.read_text().splitlines(True) if not l.startswith("==")]
per = collections.defaultdict(lambda: [0, 0.0, 0.0])  # launches, ns, bytes
for row in csv.DictReader(io.StringIO("".join(lines))):
    name = row["Kernel Name"].split("(")[0].split("<")[0][-60:]
    v = float(row["Metric Value"].replace(",", ""))
    unit, metric = row["Metric Unit"], row["Metric Name"]
    if metric == "gpu__time_duration.sum":
        per[name][0] += 1
        per[name][1] += v * {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
    elif metric.startswith("dram__bytes"):
        per[name][2] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
tot_ns = sum(v[1] for v in per.values())
tot_b = sum(v[2] for v in per.values())
gemm = [v for k, v in per.items() if "gemm_tcgen05" in k]
g_n, g_ns, g_b = sum(v[0] for v in gemm), sum(v[1] for v in gemm), sum(v[2] for v in gemm)
out = {workload: {
    "source": f"profiles/{tag}_launches.csv.gz (ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
              "dram__bytes_write.sum --clock-control none over scripts/gpu_train_step_profile.py: one fwd+bwd step + fused optimizer step)",
    "gemm_dram_bytes_per_launch": g_b / max(g_n, 1), "gemm_launches": g_n,
    "gemm_share_of_device_time": g_ns / tot_ns, "device_ms_under_ncu": tot_ns / 1e6,
    "step_dram_gb": tot_b / 1e9,
    "per_kernel": {k: {"launches": v[0], "ms": round(v[1] / 1e6, 4), "dram_gb": round(v[2] / 1e9, 4),
                       "gbs": round(v[2] / max(v[1], 1e-9), 1)} for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:40]}}}
(Path(__file__).resolve().parent.parent / "profiles" / f"{tag}_traffic.json").write_text(json.dumps(out, indent=1))
print(json.dumps({k: v for k, v in out[workload].items() if k != "per_kernel"}, indent=1))
