"""Per-shape table of the fvit_gemm launches of one step from `bench.py --profile-out <json>` (the committed launch
table): launches, total ms and algorithmic TFLOP/s per (phase, shape, epilogue) group, largest first.

    python scripts/gemm_shapes.py profiles/r02k_fv4_train_launch_table.json > profiles/r02k_gemm_shapes.txt
"""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
rows = [r for r in d["launches"] if r["name"] == "fvit_gemm"]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    k = (r.get("phase", ""), r.get("shape", ""))
    agg[k][0] += 1
    agg[k][1] += r["ms"]
    agg[k][2] += r["flops"]
tot_ms = sum(v[1] for v in agg.values())
tot_fl = sum(v[2] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} fvit_gemm launches, {tot_ms:.2f} ms, {tot_fl / tot_ms / 1e9:.0f} TF/s algorithmic overall")
print("# shape = m n k x taps, split-K, activation code, flags (A/B MN-major operands, r resid, m row map, 3 fp32 out, h fp16 out, "
      "s BN statistics, p pre-activation / gelu' out, x aux in)")
print(f"# {'phase':5s} {'shape':54s} {'n':>4s} {'ms':>8s} {'TF/s':>6s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[0]:5s} {k[1]:54s} {v[0]:4d} {v[1]:8.3f} {v[2] / v[1] / 1e9:6.0f} {v[1] / tot_ms:6.3f}")
