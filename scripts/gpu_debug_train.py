"""GPU debug aid: train-mode forward + backward of one golden case; prints logits / loss / BN-stat errors and
the per-parameter gradient errors (worst first) against the fp64 reference goldens."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fastervit_b200 as F  # noqa: E402
from oracle import fastervit_oracle as O  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "tiny_a"
g = torch.load(ROOT / "tests" / "golden" / f"{case}.pt", weights_only=False)
tr = g["train"]
model = F.create_model(g["entry"], drop_path_rate=0.0, **g["kwargs"])
O.synth_fill_(model.state_dict(), g["seeds"]["w"])
model = model.cuda().train()
x = O.synth_input(tr["batch"], g["cfg"]["resolution"], g["seeds"]["x"] + 100, torch.float32).cuda()
logits = model(x)
loss = torch.nn.functional.cross_entropy(logits, tr["target"].cuda())
loss.backward()
torch.cuda.synchronize()
ref = tr["logits"].cuda()
print("train logits max-rel err:", ((logits.double() - ref).abs().max() / ref.abs().max()).item(), " loss", loss.item(), "ref", tr["loss"])
sd = model.state_dict()
worst = 0
for k, want in tr["bn_after"].items():
    f = sd[k].flatten()
    stride = max(1, (f.numel() + 511) // 512)
    d = (f[::stride].float().cpu() - want["sample"]).abs().max().item() / max(want["amax"], 1e-12)
    worst = max(worst, d)
print("BN running stats worst rel err:", worst)
gmax = max(w["amax"] for w in tr["grads"].values())
floor = 1e-3 * gmax
rows = []
for k, p in model.named_parameters():
    want = tr["grads"].get(k)
    if want is None:
        continue
    if p.grad is None:
        rows.append((float("inf"), k, 0.0, want["l2"]))
        continue
    f = p.grad.flatten()
    stride = max(1, (f.numel() + 511) // 512)
    smp = f[::stride].float().cpu()
    d = (smp - want["sample"]).abs().max().item() / max(want["amax"], floor)
    rows.append((d, k, p.grad.double().norm().item(), want["l2"]))
print(f"{len(rows)} parameter gradients (sample max err / max(amax, 1e-3*global amax), |g| got, |g| ref), model order:")
for r in rows:
    flag = "" if r[0] < 2e-2 else ("  <-- BAD" if r[0] > 0.1 else "  <- meh")
    print(f"  {r[0]:.3e}  {r[1]:62s} {r[2]:.4e} {r[3]:.4e}{flag}")
ok = sum(1 for r in rows if r[0] < 2e-2)
print(f"{ok}/{len(rows)} within 2e-2")
