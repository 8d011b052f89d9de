"""GPU micro-benchmark of the memory-bound training kernels at faster_vit_4 (batch 128) shapes: prints the time
and the algorithmic HBM bandwidth of each launch. Run under ncu to look at one of them:
    ncu --set full -k regex:bn_bwd -c 4 python scripts/gpu_ops_micro.py bn_bwd"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fastervit_b200 import lib as L  # noqa: E402

which = set(sys.argv[1:])
dev = "cuda"
B, Hh, C0 = 128, 56, 196
Hp = Hh + 2
rows_pad = B * Hp * Hp
b_i, y_i, x_i = torch.meshgrid(torch.arange(B), torch.arange(Hh), torch.arange(Hh), indexing="ij")
pix = (b_i * Hp * Hp + (y_i + 1) * Hp + x_i + 1).reshape(-1).to(torch.int32).to(dev)
npix = pix.numel()
ld = 200


def timed(name, nbytes, fn, reps=5):
    if which and name not in which:
        return
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(reps):
        big.zero_()  # flush L2
        ev[0].record()
        fn()
        ev[1].record()
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:28s} {t * 1000:8.1f} us   {nbytes / t / 1e6:8.1f} GB/s algorithmic")


g32 = torch.randn(rows_pad, C0, device=dev)
raw16 = torch.randn(rows_pad, ld, device=dev).half()
out16 = torch.zeros(rows_pad, ld, device=dev, dtype=torch.float16)
vecs = [torch.rand(C0, device=dev) + 0.5 for _ in range(6)]
s1, s2, dw, db = (torch.zeros(C0, device=dev) for _ in range(4))
one = torch.ones(1, device=dev)
p = lambda t: t.data_ptr()
timed("bn_bwd", npix * C0 * (4 + 2) * 2 + npix * C0 * 2,
      lambda: L.call("fvit_bn_bwd", p(g32), 0, C0, p(pix), p(raw16), ld, p(pix), npix, C0, p(vecs[0]), p(vecs[1]), p(vecs[2]),
                     p(vecs[3]), L.ACT_NONE, None, p(s1), p(s2), p(one), p(out16), ld, p(pix), p(dw), p(db), None))
dy16 = torch.randn(rows_pad, ld, device=dev).half()
timed("bn_bwd_f16", npix * C0 * (2 + 2) * 2 + npix * C0 * 2,
      lambda: L.call("fvit_bn_bwd", p(dy16), 1, ld, p(pix), p(raw16), ld, p(pix), npix, C0, p(vecs[0]), p(vecs[1]), p(vecs[2]),
                     p(vecs[3]), L.ACT_NONE, None, p(s1), p(s2), p(one), p(out16), ld, p(pix), p(dw), p(db), None))
o32 = torch.zeros(rows_pad, C0, device=dev)
timed("affine_rows", npix * C0 * (2 + 4 + 4 + 2),
      lambda: L.call("fvit_affine_rows", p(raw16), ld, p(pix), npix, C0, p(vecs[0]), p(vecs[1]), L.ACT_NONE, p(g32), C0, p(o32), C0,
                     p(out16), ld, None))
timed("affine_rows_gelu16", npix * C0 * (2 + 2),
      lambda: L.call("fvit_affine_rows", p(raw16), ld, p(pix), npix, C0, p(vecs[0]), p(vecs[1]), L.ACT_GELU, None, 0, None, 0,
                     p(out16), ld, None))
# token levels (level 2: 27136 token rows, C = 784, hidden 3136)
R, C2, Hd = 27136, 784, 3136
a16 = torch.randn(R, Hd, device=dev).half()
outv = torch.zeros(Hd, device=dev)
timed("colsum_3136", R * Hd * 2, lambda: L.call("fvit_colsum", p(a16), 1, Hd, None, None, 0, R, Hd, None, p(one), p(outv), None))
a16b = torch.randn(R, C2, device=dev).half()
timed("colsum_784", R * C2 * 2, lambda: L.call("fvit_colsum", p(a16b), 1, C2, None, None, 0, R, C2, None, p(one), p(outv), None))
g2 = torch.randn(R, C2, device=dev)
timed("colsum_784_f32", R * C2 * 4, lambda: L.call("fvit_colsum", p(g2), 0, C2, None, None, 0, R, C2, None, p(one), p(outv), None))
xh = torch.randn(R, C2, device=dev).half()
rs = torch.rand(R, device=dev) + 0.5
gam = torch.rand(C2, device=dev) + 0.5
dg, dbt = torch.zeros(C2, device=dev), torch.zeros(C2, device=dev)
timed("ln_bwd_784", R * C2 * (2 + 2 + 4 + 4) + R * C2 * 4,
      lambda: L.call("fvit_ln_bwd", p(a16b), C2, None, p(xh), C2, p(rs), p(gam), R, C2, p(g2), C2, None, 1, 0, p(one), p(dg), p(dbt)))
x32 = torch.randn(R, C2, device=dev)
timed("cast_scale_784", R * C2 * 6,
      lambda: L.call("fvit_cast_scale_f16", p(x32), C2, None, R, C2, None, p(one), p(a16b), C2, None))
# positional MLP backward: P = 49 window positions, D = 784; and the bias table P = 169.., D = heads
for P, D in ((49, 784), (196, 784), (169, 16)):
    coords = torch.randn(P, 2, device=dev)
    w1 = torch.randn(D, 512, device=dev)
    hid = torch.rand(P, 512, device=dev)
    dout = torch.randn(P, D, device=dev)
    dw0, db0, dw1 = torch.zeros(512, 2, device=dev), torch.zeros(512, device=dev), torch.zeros(D, 512, device=dev)
    timed(f"cpb_mlp_bwd", (D * 512 * 3 + P * D + P * 512) * 4,
          lambda: L.call("fvit_cpb_mlp_bwd", p(coords), P, p(w1), p(hid), p(dout), D, p(one), p(dw0), p(db0), p(dw1)))
