"""Turn gpurun_out/ ncu artefacts into the small, tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [--launches launches.csv] [--rep name.ncu-rep ...]

 * launches csv (ncu --metrics gpu__time_duration.sum): per-kernel launch count / total time / share;
 * each .ncu-rep (ncu --set full): per captured launch the duration, DRAM bytes, tensor-pipe %, L2/L1
   throughput %, registers, issue-stall mix — read with `ncu -i ... --page raw --csv`.
"""
import collections
import csv
import gzip
import io
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path: Path, tag: str) -> None:
    lines = [l for l in path.read_text().splitlines(True) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])   # launches, us, dram bytes
    for row in csv.DictReader(io.StringIO("".join(lines))):
        name = row["Kernel Name"].split("(")[0][-70:]
        metric = row.get("Metric Name")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        if metric == "gpu__time_duration.sum":
            us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
            agg[name][0] += 1
            agg[name][1] += us
        elif metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            agg[name][2] += v * mult
    tot = sum(v[1] for k, v in agg.items() if "spin_kernel" not in k)
    out = [f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --clock-control none` ({path.name})",
           "# cold-cache, serialised launches: compare SHARES with bench.py's live CUDA-event table, not absolutes",
           f"# total (excluding torch's spin kernel): {tot / 1e3:.3f} ms", ""]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = v[1] / tot if "spin_kernel" not in k else float("nan")
        dram = f"  dram {v[2] / v[0] / 1e6:9.2f} MB/launch {v[2] / max(v[1], 1e-9) / 1e3:8.1f} GB/s" if v[2] else ""
        out.append(f"{k:72s} {v[0]:6d} launches {v[1] / 1e3:10.3f} ms  share {share:.4f}{dram}")
    (OUT / f"{tag}_launches_summary.txt").write_text("\n".join(out) + "\n")
    with gzip.open(OUT / f"{tag}_launches.csv.gz", "wt") as f:
        f.write("".join(lines))


def rep(path: Path, tag: str) -> None:
    res = subprocess.run(["ncu", "-i", str(path), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(res.stdout)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# `ncu --set full --clock-control none` capture {path.name}: selected raw metrics per captured launch", ""]
    for n, r in enumerate(rows[2:]):
        out.append(f"## launch {n}: {r[idx['Kernel Name']][:100]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in KEYS:
            if k in idx:
                out.append(f"  {k:90s} {r[idx[k]]} {units[idx[k]]}")
        out.append("")
    (OUT / f"{tag}_{path.stem}.txt").write_text("\n".join(out))


if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    tag = sys.argv[1]
    args = sys.argv[2:]
    i = 0
    while i < len(args):
        if args[i] == "--launches":
            launches(Path(args[i + 1]), tag)
        elif args[i] == "--rep":
            rep(Path(args[i + 1]), tag)
        i += 2
