"""GPU debug aid: does a backward pass disturb the next forward? Runs fwd, fwd, bwd, fwd on one golden case and
diffs the logits plus every named plan buffer written by the forward list."""
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


def snap(plan):
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in plan.bufs.named.items() if v.dtype in (torch.float16, torch.float32)}


l1 = model(x)
plan = next(iter(model._get_engine().plans.values()))
s1 = snap(plan)
l1b = model(x)
s1b = snap(plan)
print("fwd, fwd: logits diff", (l1b - l1).abs().max().item(), "of", l1.abs().max().item())
loss = torch.nn.functional.cross_entropy(l1b, tr["target"].cuda())
loss.backward()
l2 = model(x)
s2 = snap(plan)
print("fwd, bwd, fwd: logits diff", (l2 - l1).abs().max().item())


def diff(a, b, tag):
    rows = []
    for k in a:
        d = (a[k].float() - b[k].float()).abs().max().item()
        m = a[k].float().abs().max().item()
        if d > 1e-5 * max(m, 1e-6):
            rows.append((k, d, m))
    print(f"{tag}: {len(rows)} buffers differ")
    for r in rows[:60]:
        print("   %-50s diff %.3e  amax %.3e" % r)


diff(s1, s1b, "fwd vs fwd")
diff(s1, s2, "fwd vs fwd-after-bwd")
