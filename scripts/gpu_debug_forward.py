"""GPU debug aid: run one golden case through the engine and print, level by level, the relative error
of the engine's persistent buffers against the CPU oracle's captured activations (fp32)."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fastervit_b200 as F  # noqa: E402
from oracle import fastervit_oracle as O  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "tiny_a"
g = torch.load(ROOT / "tests" / "golden" / f"{case}.pt", weights_only=False)
model = F.create_model(g["entry"], drop_path_rate=0.0, **g["kwargs"]).eval()
sd = model.state_dict()
O.synth_fill_(sd, g["seeds"]["w"])
B = g["eval"]["batch"]
x = O.synth_input(B, g["cfg"]["resolution"], g["seeds"]["x"], torch.float32)
cap = {}
with torch.no_grad():
    ref_logits = O.forward({k: v.clone() for k, v in sd.items()}, g["cfg"], x, capture=cap)
model = model.cuda()
t0 = time.time()
with torch.no_grad():
    out = model(x.cuda())
torch.cuda.synchronize()
print(f"forward ok in {time.time() - t0:.2f}s; launches so far:", F.lib.launch_count() if hasattr(F, 'lib') else '')
plan = next(iter(model._engine.plans.values()))


def rel(a, b):
    return ((a.float().cpu() - b.float()).abs().max() / b.float().abs().max()).item()


nb = plan.bufs.named
for i, level in enumerate(model.levels):
    want = cap[f"levels.{i}.out"]  # [B, C, H, W]
    Bc, Cc, Hc, Wc = want.shape
    if level.conv:
        got = nb[f"l{i}.x32"].view(Bc, Hc + 2, Wc + 2, Cc)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
        brd = nb[f"l{i}.x32"].view(Bc, Hc + 2, Wc + 2, Cc)
        border_max = max(brd[:, 0].abs().max().item(), brd[:, -1].abs().max().item(),
                         brd[:, :, 0].abs().max().item(), brd[:, :, -1].abs().max().item())
        print(f"level {i} (conv) out rel err {rel(got, want):.3e}  border max {border_max:.1e}")
    else:
        rows = nb[f"l{i}.crop_map"].long()
        got = nb[f"l{i}.xs"][rows].view(Bc, Hc, Wc, Cc).permute(0, 3, 1, 2)
        print(f"level {i} (tok)  out rel err {rel(got, want):.3e}")
    if i == 0:
        pass
print("logits rel err vs oracle fp32:", rel(out, ref_logits))
print("logits rel err vs golden fp64:", rel(out, g["eval"]["logits"]))
