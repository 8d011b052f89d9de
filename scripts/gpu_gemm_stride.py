"""GPU micro-benchmark: does the row stride of a K-major / MN-major fp16 operand matter (L2 slice camping)?
Times fvit_gemm for the same logical GEMM with the operand stored at different leading dimensions."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fastervit_b200 import lib as L  # noqa: E402

dev = "cuda"


def bench(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def case(m, n, k, tag):
    print(f"--- {tag}: m={m} n={n} k={k}")
    for pad in (0, 8, 64, 128):
        A = torch.randn(m, k + pad, device=dev).half()[:, :k]          # K-major A, lda = k + pad
        W = torch.randn(n, k + 8, device=dev).half()[:, :k]
        out = torch.empty(m, n, device=dev, dtype=torch.float16)
        t = bench(lambda: L.gemm(A, W, out_f16=out))
        print(f"  A K-major   lda={k + pad:6d} ({(k + pad) * 2:6d} B): {t * 1000:7.1f} us  {2 * m * n * k / t / 1e9:7.1f} TF/s")
    for pad in (0, 8, 64):
        # wgrad-like: out[n_out=k? ...]: A MN-major [rows=m, cols=k+pad] as dz (M = k), B MN-major x [m, n]
        dz = torch.randn(m, k + pad, device=dev).half()[:, :k]
        x = torch.randn(m, n + 8, device=dev).half()[:, :n]
        o32 = torch.zeros(k, n, device=dev)
        t = bench(lambda: L.gemm(dz, x, a_mn=True, b_mn=True, split_k=3, out_f32=o32))
        print(f"  A MN-major  lda={k + pad:6d}: wgrad [{k}x{n}] over {m} rows {t * 1000:7.1f} us  {2 * m * n * k / t / 1e9:7.1f} TF/s")


case(27136, 784, 3072, "qkv dgrad (fv4 L2)")
case(27136, 784, 1024, "proj fwd (fv4 L2)")
case(54272, 256, 1024, "fc2 fwd (fv0 L2)")
case(16384, 512, 2048, "fc2 fwd (fv0 L3)")
