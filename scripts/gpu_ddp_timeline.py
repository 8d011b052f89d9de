"""Per-launch timeline of the fv4 backward pass with and without the gradient all-reduce in flight (torchrun, N >= 2).

For rank 0 it prints, per variant, the total backward device time and the launches whose duration changed most
against the no-all-reduce run: evidence of where the data-parallel step loses time (NCCL CTAs vs the persistent
148-CTA GEMM grids) and of what the SM cap in the all-reduce shadow (engine_train._allreduce_shadow) buys.
    torchrun --nproc-per-node 2 scripts/gpu_ddp_timeline.py [out.json]
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("NCCL_MAX_CTAS", "16")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import fastervit_b200 as F  # noqa: E402
from fastervit_b200.engine_train import GradBucketReducer  # noqa: E402

torch.manual_seed(0)
model = F.create_model("faster_vit_4_224", drop_path_rate=0.0).to(dev).train()
B = 128
x = torch.randn(B, 3, 224, 224, device=dev)
tgt = torch.randint(0, 1000, (B,), device=dev)
for _ in range(2):   # builds the plan, warms everything up (no all-reduce yet)
    torch.nn.functional.cross_entropy(model(x), tgt).backward()
    model.zero_grad(set_to_none=True)
plan = next(p for p in model._engine.plans.values() if p.training)
dl = torch.full((B, 1000), 1.0 / B, device=dev)


def timed_backward(mode: str):
    """mode: 'local' (no all-reduce), 'ar' (all-reduce, no SM cap), 'ar_cap' (all-reduce + shadow cap)"""
    plan.run_forward(x)
    plan._bwd_start(dl)
    red = None
    if mode != "local":
        red = GradBucketReducer(plan.gflat, None)
        plan._ar_active = red
        plan._ar_shadow = plan._allreduce_shadow(world) if mode == "ar_cap" else {}
    ops = plan.bwd_ops
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda._sleep(int(6e8))
    evs[0].record()
    st = plan.lib
    for i, op in enumerate(ops):
        if red is not None:
            st.fvit_set_sm_limit(plan._ar_shadow.get(i, 0))
        plan.run_ops([op], None) if op[0] != "bucket" else (red.reduce(*op[1]) if red is not None else None)
        evs[i + 1].record()
    st.fvit_set_sm_limit(0)
    if red is not None:
        red.finish(plan.grad_buckets)
        plan._ar_active = None
    torch.cuda.synchronize()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(ops))]
    return ms


res = {}
for mode in ("local", "ar", "ar_cap", "local", "ar", "ar_cap"):
    ms = timed_backward(mode)
    res.setdefault(mode, []).append(ms)
if rank == 0:
    base = res["local"][-1]
    out = {"world": world, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"), "ops": len(base)}
    names = [op[2] for op in plan.bwd_ops]
    for mode in ("local", "ar", "ar_cap"):
        ms = res[mode][-1]
        delta = sorted(((ms[i] - base[i], i) for i in range(len(ms))), reverse=True)[:12]
        out[mode] = {"backward_ms": round(sum(ms), 3),
                     "largest_slowdowns_vs_local": [dict(op=i, name=names[i], ms=round(ms[i], 3), local_ms=round(base[i], 3))
                                                    for d, i in delta if d > 0.02]}
    out["bucket_ops"] = [i for i, op in enumerate(plan.bwd_ops) if op[0] == "bucket"]
    out["shadow_ops"] = len(plan._allreduce_shadow(world))
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
dist.destroy_process_group()
