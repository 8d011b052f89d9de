"""One faster_vit_4_224 training step (fwd + bwd, batch 128) followed by the fused optimizer step (clip + LAMB + EMA)
inside a cudaProfilerStart/Stop window, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
      --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/gpu_train_step_profile.py
  ncu --profile-from-start off --set full --clock-control none -k regex:optim_ -c 8 -o gpurun_out/optim \\
      python scripts/gpu_train_step_profile.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fastervit_b200 as F  # noqa: E402

entry = sys.argv[1] if len(sys.argv) > 1 else "faster_vit_4_224"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
model = F.create_model(entry, drop_path_rate=0.0).cuda().train()
big = sum(p.numel() for p in model.parameters()) > 100e6
opt = (F.FusedLamb(model, lr=5e-3, weight_decay=0.12, max_grad_norm=1.0) if big
       else F.FusedAdamW(model, lr=5e-4, weight_decay=0.05, max_grad_norm=5.0))
ema = F.FlatEma(model, decay=0.9998)
opt.attach_ema(ema, model)
x = torch.randn(batch, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (batch,), device="cuda")


def step():
    opt.zero_grad()
    loss = torch.nn.functional.cross_entropy(model(x), y)
    loss.backward()
    opt.step()
    ema.update(model)
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{entry} batch {batch}: loss {loss.item():.4f}, grad norm {opt.grad_norm.item():.4f}, "
      f"skipped {opt.found_inf_flag.item()}")
