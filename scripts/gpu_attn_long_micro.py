"""Micro-benchmark of the long-window attention backward (fvit_attn_loop_bwd_long) on the faster_vit_4_21k_384 level-2
shape at batch 32 (S = 576, 16 heads of 49 -> 64): full call, without the dbias reductions, without bias and dbias."""
import sys
import torch
sys.path.insert(0, ".")
from fastervit_b200 import lib
lib.load()
S, heads, hd, groups = 576, 16, 49, 32
if len(sys.argv) > 1:
    S, heads, hd, groups = (int(v) for v in sys.argv[1:5])
hdp = 32 if hd <= 32 else 64
g = torch.Generator(device="cuda").manual_seed(3)
qkv = torch.zeros(groups * S, 3, heads, hdp, device="cuda")
qkv[..., :hd] = torch.randn(groups * S, 3, heads, hd, device="cuda", generator=g)
qkv16 = qkv.reshape(groups * S, 3 * heads * hdp).half()
do16 = torch.randn(groups * S, heads * hdp, device="cuda", generator=g).half()
bias = torch.randn(heads, S, S, device="cuda", generator=g)
out16 = torch.zeros(groups * S, heads * hdp, device="cuda", dtype=torch.half)
lse = torch.zeros(groups * S, heads, device="cuda")
scale = hd ** -0.5
dqkv = torch.zeros(groups * S, 3 * heads * hdp, device="cuda", dtype=torch.half)
dbias = torch.zeros(heads, S, S, device="cuda")
scratch = torch.zeros(groups * S, heads * hdp, device="cuda")


def fwd(b):
    lib.call("fvit_attn_loop_fwd", qkv16.data_ptr(), qkv16.stride(0), groups, S, heads, hdp, b.data_ptr() if b is not None else None,
             scale, out16.data_ptr(), out16.stride(0), lse.data_ptr())


def bwd(b, db):
    lib.call("fvit_attn_loop_bwd_long", qkv16.data_ptr(), qkv16.stride(0), do16.data_ptr(), do16.stride(0), out16.data_ptr(),
             out16.stride(0), lse.data_ptr(), groups, S, heads, hdp, b.data_ptr() if b is not None else None, scale,
             dqkv.data_ptr(), dqkv.stride(0), db.data_ptr() if db is not None else None, scratch.data_ptr(), scratch.stride(0))


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def bwd2(b, db):   # the two-tile kernel (128 < S <= 256: every dQ tile in TMEM)
    lib.call("fvit_attn_loop_bwd", qkv16.data_ptr(), qkv16.stride(0), do16.data_ptr(), do16.stride(0), out16.data_ptr(),
             out16.stride(0), lse.data_ptr(), groups, S, heads, hdp, b.data_ptr() if b is not None else None, scale,
             dqkv.data_ptr(), dqkv.stride(0), db.data_ptr() if db is not None else None)


flops = 10.0 * groups * heads * S * S * hd
if S <= 256:
    fwd(bias)
    print(f"S={S} heads={heads} hd={hd} groups={groups} two-tile kernel [bias + dbias]: bwd {timed(lambda: bwd2(bias, dbias)):.0f} us",
          flush=True)
for tag, b, db in (("bias + dbias", bias, dbias), ("bias, no dbias", bias, None), ("no bias, no dbias", None, None)):
    fwd(b)
    t_f = timed(lambda: fwd(b))
    t = timed(lambda: bwd(b, db))
    print(f"S={S} heads={heads} hd={hd} groups={groups} [{tag}]: fwd {t_f:.0f} us, bwd {t:.0f} us = "
          f"{flops / t / 1e6:.1f} TF/s (10 S^2 hd per window-head)", flush=True)
