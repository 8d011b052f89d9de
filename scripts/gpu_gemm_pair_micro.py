"""Single-CTA vs CTA-pair (cta_group::2) fvit_gemm on the shapes that dominate the fv4 training step:
correctness against each other and CUDA-event throughput (L2 flushed between timed launches by working on
operands larger than L2 where the shape allows, otherwise rotating buffers)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fastervit_b200 import lib as L

L.load()
dev = "cuda"
shapes = [  # (name, m, n, k, a_mn, b_mn, split_k)
    ("fc1 fwd", 27136, 3136, 784, False, False, 1),
    ("fc2 fwd", 27136, 784, 3136, False, False, 1),
    ("qkv fwd", 27136, 3072, 784, False, False, 1),
    ("proj fwd", 27136, 784, 1024, False, False, 1),
    ("fc2 dgrad", 27136, 3136, 784, False, True, 1),
    ("fc1 dgrad", 27136, 784, 3136, False, True, 1),
    ("fc2 wgrad", 784, 3136, 27136, True, True, 3),
    ("fc1 wgrad", 3136, 784, 27136, True, True, 4),
    ("L3 fc1", 6272, 6272, 1568, False, False, 1),
    ("L3 fc2", 6272, 1568, 6272, False, False, 1),
    ("carrier fc1", 2048, 3136, 784, False, False, 1),
    ("big square", 8192, 8192, 8192, False, False, 1),
]
g = torch.Generator(device=dev).manual_seed(0)
for name, m, n, k, a_mn, b_mn, sk in shapes:
    A = (torch.randn(m, k, device=dev, generator=g) * 0.1).half()
    B = (torch.randn(n, k, device=dev, generator=g) * 0.1).half()
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    outs = {}
    line = f"{name:12s} m{m} n{n} k{k} {'A' if a_mn else '-'}{'B' if b_mn else '-'} sk{sk}:"
    for cg in (1, 2):
        use32 = sk > 1
        out = torch.zeros(m, n, device=dev, dtype=torch.float32 if use32 else torch.float16)
        kw = dict(out_f32=out) if use32 else dict(out_f16=out)
        for tn in (0, 256):
            try:
                for _ in range(2):
                    if use32:
                        out.zero_()
                    L.gemm(a, b, a_mn=a_mn, b_mn=b_mn, split_k=sk, cta_group=cg, tile_n=tn, **kw)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record()
                for _ in range(reps):
                    L.gemm(a, b, a_mn=a_mn, b_mn=b_mn, split_k=sk, cta_group=cg, tile_n=tn, **kw)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                line += f"  cg{cg}/tn{tn or 'auto'} {ms * 1e3:7.1f}us {2.0 * m * n * k / ms / 1e9:6.0f}TF"
                if tn == 0:
                    if use32:
                        out.zero_()
                        L.gemm(a, b, a_mn=a_mn, b_mn=b_mn, split_k=sk, cta_group=cg, tile_n=tn, **kw)
                    outs[cg] = out.float().clone()
            except Exception as e:  # noqa: BLE001
                line += f"  cg{cg}/tn{tn}: {str(e)[:80]}"
    if 1 in outs and 2 in outs:
        d = (outs[1] - outs[2]).abs().max().item() / outs[1].abs().max().item()
        line += f"  | cg1-vs-cg2 {d:.1e}"
    print(line, flush=True)
    sys.stdout.flush()
