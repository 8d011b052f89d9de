"""Does the row stride of the A operand matter? qkv data-gradient shape of fv4 level 2 (m27136 n784, B MN-major) with
K = 3072 at row strides 3072 / 3136 / 3200 elements, next to the fc2-dgrad twin (K = 3136) -- the launch table of the
training step shows 688 vs 900 TF/s for the two. Inputs rotate over 4 buffers (> L2)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fastervit_b200 import lib as L

L.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
cases = [("m27136 n784", 27136, 784, 3072, (3072, 3136, 3200, 3072 + 8)), ("m27136 n784", 27136, 784, 3136, (3136, 3200)),
         ("m6272 n1568", 6272, 1568, 6144, (6144, 6208)), ("m6272 n1568", 6272, 1568, 6272, (6272,)),
         ("qkv fwd out-stride n3072", 27136, 3072, 784, (784,))]
for name, m, n, k, lds in cases:
    for ld in lds:
        As = [(torch.randn(m, ld, device=dev, generator=g) * 0.1).half() for _ in range(4)]
        B = (torch.randn(k, n, device=dev, generator=g) * 0.1).half()   # MN-major B: [K, N]
        out = torch.zeros(m, n, device=dev, dtype=torch.float16)
        for A in As:
            L.gemm(A[:, :k], B, b_mn=True, out_f16=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 12
        e0.record()
        for i in range(reps):
            L.gemm(As[i % 4][:, :k], B, b_mn=True, out_f16=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name} k{k} lda{ld}: {ms * 1e3:7.1f} us  {2.0 * m * n * k / ms / 1e9:6.0f} TF/s", flush=True)
