"""Train-mode forward and backward emitters for the convolutional part of FasterViT (PatchEmbed,
ConvBlock levels, Downsample) — methods of TrainPlan.

BatchNorm2d runs with batch statistics: the convolution GEMM stores the raw output (fp16) and per-channel
sum / sum-of-squares, `fvit_bn_finalize` turns them into scale/shift (and updates the running statistics),
`fvit_affine_rows` normalises + activates (+ residual). The backward is the mirror image: `fvit_bn_bwd`,
the data-gradient convolution as a 9-tap GEMM over the flipped-transposed packed weights, and the
weight gradient as nine MN-major GEMMs (one per tap: dW_t = dZ^T X shifted by the tap) with split-K.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as L
from .engine import _ru


def P(t):
    return None if t is None else t.data_ptr()


def _bn_train(self, name: str, bn: nn.BatchNorm2d, count: int, ls=None) -> dict:
    """statistics accumulators + finalize launch; returns the saved vectors"""
    C = bn.num_features
    nb = self.bufs
    st = nb.new(name + ".stats", (2, C), torch.float32)
    self.fwd_zero.append(st)
    d = dict(bn=bn, st=st, sc=nb.new(name + ".scale", (C,), torch.float32), sh=nb.new(name + ".shift", (C,), torch.float32),
             mu=nb.new(name + ".mean", (C,), torch.float32), rs=nb.new(name + ".rstd", (C,), torch.float32),
             s12=nb.new(name + ".s12", (2, C), torch.float32), count=count)
    return d


def _bn_finalize(self, d: dict, ls=None) -> None:
    bn = d["bn"]
    self._op(self.ops, "fvit_bn_finalize", d["st"][0].data_ptr(), d["st"][1].data_ptr(), float(d["count"]),
             bn.weight.data_ptr(), bn.bias.data_ptr(), float(bn.eps), float(bn.momentum), bn.running_mean.data_ptr(),
             bn.running_var.data_ptr(), P(ls), d["sc"].data_ptr(), d["sh"].data_ptr(), d["mu"].data_ptr(),
             d["rs"].data_ptr(), bn.num_features)


def _emit_conv_part_train(self) -> dict:
    m, B, nb = self.model, self.B, self.bufs
    cfg = m.cfg
    dim, in_dim = cfg["dim"], cfg["in_dim"]
    H1, W1 = (self.H + 1) // 2, (self.W + 1) // 2
    H0, W0 = (H1 + 1) // 2, (W1 + 1) // 2
    ld_in = _ru(in_dim, 8)
    pl_rows = B * (H0 + 1) * (W0 + 1)
    npix1 = B * H1 * W1
    bi, yi, xi = torch.meshgrid(torch.arange(B), torch.arange(H1), torch.arange(W1), indexing="ij")
    plane = (yi & 1) * 2 + (xi & 1)
    stem_map = nb.i32("stem.map", (plane * pl_rows + bi * (H0 + 1) * (W0 + 1) + ((yi >> 1) + 1) * (W0 + 1)
                                   + (xi >> 1) + 1).reshape(-1))
    pe = m.patch_embed.conv_down
    st = dict(H1=H1, W1=W1, H0=H0, W0=W0, pl_rows=pl_rows, ld_in=ld_in, npix1=npix1, stem_map=stem_map)
    # conv1: im2col + GEMM -> raw1 (plane layout) + statistics
    col16 = nb.new("stem.col16", (npix1, self.STEM_LD), torch.float16)
    self._x_args.append(dict(B=B, cin=cfg["in_chans"], H=self.H, W=self.W, out=col16.data_ptr(), ldo=self.STEM_LD))
    self.ops.append(("im2col", len(self._x_args) - 1, "fvit_stem_im2col"))
    w1 = nb.new("stem.conv1.w16", (in_dim, self.STEM_LD), torch.float16)
    self._op(self.prep_ops, "fvit_cast_pad_f16", pe[0].weight.data_ptr(), 27, w1.data_ptr(), self.STEM_LD, in_dim, 27, 32)
    raw1 = nb.new("stem.raw1", (4 * pl_rows, ld_in), torch.float16)
    planes = nb.new("stem.planes", (4 * pl_rows, ld_in), torch.float16)
    bn1 = _bn_train(self, "stem.bn1", pe[1], npix1)
    self._gemm(a=col16.data_ptr(), a_rows=npix1, lda=self.STEM_LD, b=w1.data_ptr(), ldb=self.STEM_LD, m=npix1, n=in_dim,
               kc=32, row_map=stem_map.data_ptr(), out_f16=raw1.data_ptr(), ld_o16=ld_in, col_sum=bn1["st"][0].data_ptr(),
               col_sumsq=bn1["st"][1].data_ptr())
    self.op_flops[len(self.ops) - 1] = 2.0 * npix1 * in_dim * 27
    _bn_finalize(self, bn1)
    self._op(self.ops, "fvit_affine_rows", raw1.data_ptr(), ld_in, stem_map.data_ptr(), npix1, in_dim, bn1["sc"].data_ptr(),
             bn1["sh"].data_ptr(), L.ACT_RELU, None, 0, None, 0, planes.data_ptr(), ld_in, None)
    # conv2 (stride 2) -> raw2 in the level-0 layout
    lvl = self._conv_level_train_buffers(0, dim, H0, W0, len(m.levels[0].blocks))
    w2, ldw2 = self._pack_conv("stem.conv2", pe[3])
    raw2 = nb.new("stem.raw2", (lvl["rows"], lvl["ld"]), torch.float16)
    bn2 = _bn_train(self, "stem.bn2", pe[4], B * H0 * W0)
    to_l0 = self._plane_to_padded_map("stem.to_l0", B, H0, W0)
    self._gemm(a=planes.data_ptr(), a_rows=pl_rows, lda=ld_in, a_planes=4, a_plane_stride=pl_rows * ld_in, b=w2.data_ptr(),
               ldb=ldw2, m=pl_rows, n=dim, kc=in_dim, taps=self._s2_taps(W0), m_alg=B * H0 * W0, row_map=to_l0.data_ptr(),
               out_f16=raw2.data_ptr(), ld_o16=lvl["ld"], col_sum=bn2["st"][0].data_ptr(), col_sumsq=bn2["st"][1].data_ptr())
    _bn_finalize(self, bn2)
    self._op(self.ops, "fvit_affine_rows", raw2.data_ptr(), lvl["ld"], lvl["pix"].data_ptr(), B * H0 * W0, dim,
             bn2["sc"].data_ptr(), bn2["sh"].data_ptr(), L.ACT_RELU, None, 0, lvl["x32"].data_ptr(), dim,
             lvl["x16s"][0].data_ptr(), lvl["ld"], None)
    st.update(col16=col16, w1=w1, raw1=raw1, planes=planes, bn1=bn1, w2=w2, ldw2=ldw2, raw2=raw2, bn2=bn2, to_l0=to_l0)
    self.stem_sv = st
    # conv levels
    self.conv_lv = []
    Hc, Wc, Cc = H0, W0, dim
    prev = None
    for i, level in enumerate(m.levels):
        if not level.conv:
            break
        if i > 0:
            Hc, Wc, Cc = (Hc + 1) // 2, (Wc + 1) // 2, Cc * 2
            lvl = self._conv_level_train_buffers(i, Cc, Hc, Wc, len(level.blocks))
            self._emit_downsample_train(i - 1, prev, lvl)
        self._emit_conv_blocks_train(i, level, lvl)
        lvl["kind"] = "conv"
        self.conv_lv.append(lvl)
        prev = lvl
    self.conv_out = prev
    return prev


def _conv_level_train_buffers(self, i: int, Cc: int, Hc: int, Wc: int, depth: int) -> dict:
    B, nb = self.B, self.bufs
    rows = B * (Hc + 2) * (Wc + 2)
    ld = _ru(Cc, 8)
    bi, yi, xi = torch.meshgrid(torch.arange(B), torch.arange(Hc + 2), torch.arange(Wc + 2), indexing="ij")
    inside = (yi >= 1) & (yi <= Hc) & (xi >= 1) & (xi <= Wc)
    ident = torch.arange(rows).view(B, Hc + 2, Wc + 2)
    d = dict(C=Cc, H=Hc, W=Wc, ld=ld, rows=rows, npix=B * Hc * Wc, kind="conv",
             interior=nb.i32(f"l{i}.interior", torch.where(inside, ident, torch.full_like(ident, -1)).reshape(-1)),
             pix=nb.i32(f"l{i}.pix", ident[:, 1:-1, 1:-1].reshape(-1)),
             x32=nb.new(f"l{i}.x32", (rows, Cc), torch.float32),
             x16s=[nb.new(f"l{i}.x16.{j}", (rows, ld), torch.float16) for j in range(depth + 1)],
             dzA=nb.new(f"l{i}.dzA", (rows, ld), torch.float16), dzB=nb.new(f"l{i}.dzB", (rows, ld), torch.float16),
             dyA=nb.new(f"l{i}.dyA", (rows, ld), torch.float16),
             g=nb.new(f"l{i}.g", (rows + 1, Cc), torch.float32), zero_row=rows)
    d["x16"] = d["x16s"][-1]  # level output operand (input of the downsample LayerNorm is x32)
    return d


def _emit_conv_blocks_train(self, i: int, level, lv: dict) -> None:
    B, nb = self.B, self.bufs
    Cc, Wp, rows, ld, npix = lv["C"], lv["W"] + 2, lv["rows"], lv["ld"], lv["npix"]
    taps = [((dy - 1) * Wp + (dx - 1), 0) for dy in range(3) for dx in range(3)]
    lv["taps"] = taps
    lv["blocks_sv"] = []
    for j, blk in enumerate(level.blocks):
        if hasattr(blk, "gamma"):
            raise L.FvitError("layer_scale_conv is not supported by the training kernels (unused by every shipped config)")
        nm = f"l{i}.b{j}"
        rate = self.model.drop_path_rates[self._block_counter]
        self._block_counter += 1
        rs = self._drop_buf(f"levels.{i}.blocks.{j}", rate, B, (lv["H"] + 2) * (lv["W"] + 2))
        w1, ld1 = self._pack_conv(nm + ".conv1", blk.conv1)
        w2, ld2 = self._pack_conv(nm + ".conv2", blk.conv2)
        rawA = nb.new(nm + ".rawA", (rows, ld), torch.float16)
        rawB = nb.new(nm + ".rawB", (rows, ld), torch.float16)
        h16 = nb.new(nm + ".h16", (rows, ld), torch.float16)   # kept for the conv2 weight gradient (zero borders)
        bnA = _bn_train(self, nm + ".bn1", blk.norm1, npix)
        bnB = _bn_train(self, nm + ".bn2", blk.norm2, npix)
        xin, xout = lv["x16s"][j], lv["x16s"][j + 1]
        self._gemm(a=xin.data_ptr(), a_rows=rows, lda=ld, b=w1.data_ptr(), ldb=ld1, m=rows, n=Cc, kc=Cc, taps=taps,
                   m_alg=npix, col_shift=blk.conv1.bias.data_ptr(), row_map=lv["interior"].data_ptr(),
                   out_f16=rawA.data_ptr(), ld_o16=ld, col_sum=bnA["st"][0].data_ptr(), col_sumsq=bnA["st"][1].data_ptr())
        _bn_finalize(self, bnA)
        self._op(self.ops, "fvit_affine_rows", rawA.data_ptr(), ld, lv["pix"].data_ptr(), npix, Cc, bnA["sc"].data_ptr(),
                 bnA["sh"].data_ptr(), L.ACT_GELU, None, 0, None, 0, h16.data_ptr(), ld, None)
        self._gemm(a=h16.data_ptr(), a_rows=rows, lda=ld, b=w2.data_ptr(), ldb=ld2, m=rows, n=Cc, kc=Cc, taps=taps,
                   m_alg=npix, col_shift=blk.conv2.bias.data_ptr(), row_map=lv["interior"].data_ptr(),
                   out_f16=rawB.data_ptr(), ld_o16=ld, col_sum=bnB["st"][0].data_ptr(), col_sumsq=bnB["st"][1].data_ptr())
        _bn_finalize(self, bnB)
        self._op(self.ops, "fvit_affine_rows", rawB.data_ptr(), ld, lv["pix"].data_ptr(), npix, Cc, bnB["sc"].data_ptr(),
                 bnB["sh"].data_ptr(), L.ACT_NONE, lv["x32"].data_ptr(), Cc, lv["x32"].data_ptr(), Cc, xout.data_ptr(), ld, rs)
        # data-gradient operands: flipped / transposed packed weights
        wT1 = nb.new(nm + ".conv1.wT16", (Cc, 9 * _ru(Cc, 64)), torch.float16)
        wT2 = nb.new(nm + ".conv2.wT16", (Cc, 9 * _ru(Cc, 64)), torch.float16)
        self._op(self.prep_ops, "fvit_pack_conv3x3_f16", blk.conv1.weight.data_ptr(), wT1.data_ptr(), Cc, Cc, _ru(Cc, 64), 1)
        self._op(self.prep_ops, "fvit_pack_conv3x3_f16", blk.conv2.weight.data_ptr(), wT2.data_ptr(), Cc, Cc, _ru(Cc, 64), 1)
        lv["blocks_sv"].append(dict(blk=blk, rawA=rawA, rawB=rawB, h16=h16, bnA=bnA, bnB=bnB, xin=xin, wT1=wT1, wT2=wT2,
                                    rs=rs))


def _emit_downsample_train(self, i: int, src: dict, dst: dict) -> None:
    """Downsample i in training: LayerNorm saves xhat / rstd; the conv also emits a fp16 copy of the next
    level's input (the conv-level operand, or the token level's x0 for the tokenizer weight gradient)."""
    B, nb = self.B, self.bufs
    ds_mod = self.model.levels[i].downsample
    Cs, Hs, Ws = src["C"], src["H"], src["W"]
    Ho, Wo = (Hs + 1) // 2, (Ws + 1) // 2
    ld = _ru(Cs, 8)
    pl_rows = B * (Ho + 1) * (Wo + 1)
    npix = B * Hs * Ws
    planes = nb.new(f"ds{i}.planes", (4 * pl_rows, ld), torch.float16)
    bi, yi, xi = torch.meshgrid(torch.arange(B), torch.arange(Hs), torch.arange(Ws), indexing="ij")
    omap = nb.i32(f"ds{i}.omap", (((yi & 1) * 2 + (xi & 1)) * pl_rows + bi * (Ho + 1) * (Wo + 1)
                                  + ((yi >> 1) + 1) * (Wo + 1) + (xi >> 1) + 1).reshape(-1))
    src_rows = src["pix"] if src["kind"] == "conv" else src["crop_map"]
    src_x = src["x32"] if src["kind"] == "conv" else src["xs"]
    xh = nb.new(f"ds{i}.xhat", (npix, Cs), torch.float16)
    rs = nb.new(f"ds{i}.rstd", (npix,), torch.float32)
    mu = nb.new(f"ds{i}.mean", (npix,), torch.float32)
    self._op(self.ops, "fvit_ln_fwd", src_x.data_ptr(), Cs, src_rows.data_ptr(), npix, Cs, None, 1, 0, None, 0,
             ds_mod.norm.weight.data_ptr(), ds_mod.norm.bias.data_ptr(), float(ds_mod.norm.eps), planes.data_ptr(), ld,
             omap.data_ptr(), mu.data_ptr(), rs.data_ptr(), xh.data_ptr(), Cs)
    w16, ldw = self._pack_conv(f"ds{i}.conv", ds_mod.reduction[0])
    taps = self._s2_taps(Wo)
    if dst.get("kind", "tok") == "conv" or "x16s" in dst:
        rmap_t = torch.full((B, Ho + 1, Wo + 1), -1, dtype=torch.int64)
        bb, aa, cc = torch.meshgrid(torch.arange(B), torch.arange(Ho), torch.arange(Wo), indexing="ij")
        rmap_t[:, 1:, 1:] = bb * (Ho + 2) * (Wo + 2) + (aa + 1) * (Wo + 2) + cc + 1
        rmap = nb.i32(f"ds{i}.rmap", rmap_t.reshape(-1))
        self._gemm(a=planes.data_ptr(), a_rows=pl_rows, lda=ld, a_planes=4, a_plane_stride=pl_rows * ld, b=w16.data_ptr(),
                   ldb=ldw, m=pl_rows, n=2 * Cs, kc=Cs, taps=taps, m_alg=B * Ho * Wo, row_map=rmap.data_ptr(),
                   out_f32=dst["x32"].data_ptr(), ld_o32=dst["C"], out_f16=dst["x16s"][0].data_ptr(), ld_o16=dst["ld"])
    else:
        pm = dst["pix_map_host"]
        rmap_t = torch.full((B, Ho + 1, Wo + 1), -1, dtype=torch.int64)
        rmap_t[:, 1:, 1:] = pm[:, :Ho, :Wo]
        rmap = nb.i32(f"ds{i}.rmap", rmap_t.reshape(-1))
        dst["x0_16"] = nb.new(f"ds{i}.x0_16", (dst["xs"].shape[0], 2 * Cs), torch.float16)
        self._gemm(a=planes.data_ptr(), a_rows=pl_rows, lda=ld, a_planes=4, a_plane_stride=pl_rows * ld, b=w16.data_ptr(),
                   ldb=ldw, m=pl_rows, n=2 * Cs, kc=Cs, taps=taps, m_alg=B * Ho * Wo, row_map=rmap.data_ptr(),
                   out_f32=dst["xs"].data_ptr(), ld_o32=dst["C"], out_f16=dst["x0_16"].data_ptr(), ld_o16=2 * Cs)
    # backward-side maps / packed operands
    zero_row = dst["zero_row"]
    rmap_g = nb.i32(f"ds{i}.rmap_g", torch.where(rmap_t >= 0, rmap_t, torch.full_like(rmap_t, zero_row)).reshape(-1))
    plane_taps = {p: [t for t in range(9) if taps[t][1] == p] for p in range(4)}
    co_pad = _ru(2 * Cs, 64)
    wT = {}
    for p, tl_ in plane_taps.items():
        buf = nb.new(f"ds{i}.wT.{p}", (Cs, len(tl_) * co_pad), torch.float16)
        t9 = tl_ + [0] * (9 - len(tl_))
        self._op(self.prep_ops, "fvit_pack_conv3x3_taps_f16", ds_mod.reduction[0].weight.data_ptr(), buf.data_ptr(), 2 * Cs,
                 Cs, co_pad, 1, len(tl_), *t9)
        wT[p] = buf
    dst["ds"] = dict(i=i, mod=ds_mod, Cs=Cs, Hs=Hs, Ws=Ws, Ho=Ho, Wo=Wo, ld=ld, pl_rows=pl_rows, npix=npix, planes=planes,
                     omap=omap, src_rows=src_rows, xh=xh, rs=rs, taps=taps, rmap_g=rmap_g, plane_taps=plane_taps, wT=wT,
                     co_pad=co_pad, dY=nb.new(f"ds{i}.dY16", (pl_rows, 2 * Cs), torch.float16),
                     dplanes=nb.new(f"ds{i}.dplanes16", (4 * pl_rows, ld), torch.float16))


# ====================================================================================== backward
def _conv_wgrad(self, *, dz, lddz, x, ldx, x_rows, rows, cout, cin, shifts, conv_weight, name: str) -> None:
    """dW[co][ci][t] = sum_q dz[q][co] * x[q + shift_t][ci]: ONE MN-major split-K GEMM whose N dimension
    enumerates (tap, ci) -- every dz tile is fetched once for the nine shifted views of x -- plus a repack."""
    nt = len(shifts)
    scr = ("scr", self._scratch(name + ".dWtaps", nt * cout * cin))
    sk = self._split_k(cout, cin, rows, nt)
    n0 = len(self.bwd_ops)
    self._bgemm(a=dz, a_rows=rows, lda=lddz, a_mn=True, b=x, b_rows=x_rows, ldb=ldx, b_mn=True, b_taps=list(shifts), m=cout,
                n=cin, kc=rows, split_k=sk, alpha_ptr=("scal", 1), out_f32=scr, ld_o32=nt * cin,
                flops=2.0 * rows * cout * cin * nt)
    self._op(self.bwd_ops, "fvit_unpack_conv_grad", scr, nt * cin, self.G(conv_weight), cout, cin)
    self._side_from(n0, dz)   # (its own zeroed scratch slice in, a parameter gradient out: a leaf pair)


def _emit_downsample_bwd(self, ds: dict, src: dict, dst: dict) -> None:
    ops = self.bwd_ops
    Cs, ld, pl_rows, npix = ds["Cs"], ds["ld"], ds["pl_rows"], ds["npix"]
    mod = ds["mod"]
    # gradient of the conv output in plane space (border rows read the buffer's zero row)
    self._op(ops, "fvit_cast_scale_f16", dst["g"].data_ptr(), 2 * Cs, ds["rmap_g"].data_ptr(), pl_rows, 2 * Cs, None, None,
             ds["dY"].data_ptr(), 2 * Cs, None)
    shifts = [p * pl_rows + s for (s, p) in ds["taps"]]
    _conv_wgrad(self, dz=ds["dY"].data_ptr(), lddz=2 * Cs, x=ds["planes"].data_ptr(), ldx=ld, x_rows=4 * pl_rows,
                rows=pl_rows, cout=2 * Cs, cin=Cs, shifts=shifts, conv_weight=mod.reduction[0].weight, name=f"ds{ds['i']}")
    for p, tl_ in ds["plane_taps"].items():
        self._bgemm(a=ds["dY"].data_ptr(), a_rows=pl_rows, lda=2 * Cs, b=ds["wT"][p].data_ptr(), b_rows=Cs,
                    ldb=len(tl_) * ds["co_pad"], m=pl_rows, n=Cs, kc=2 * Cs, taps=[(-ds["taps"][t][0], 0) for t in tl_],
                    out_f16=ds["dplanes"].data_ptr() + 2 * p * pl_rows * ld, ld_o16=ld,
                    flops=2.0 * self.B * ds["Ho"] * ds["Wo"] * Cs * 2 * Cs * len(tl_))
    # LayerNorm2d backward: writes the source level's output gradient
    self._op(ops, "fvit_ln_bwd", ds["dplanes"].data_ptr(), ld, ds["omap"].data_ptr(), ds["xh"].data_ptr(), Cs,
             ds["rs"].data_ptr(), mod.norm.weight.data_ptr(), npix, Cs, src["g"].data_ptr(), Cs, ds["src_rows"].data_ptr(), 0, 0,
             ("scal", 1), self.G(mod.norm.weight), self.G(mod.norm.bias))


def _emit_conv_blocks_bwd(self, lv: dict) -> None:
    ops = self.bwd_ops
    Cc, rows, ld, npix = lv["C"], lv["rows"], lv["ld"], lv["npix"]
    g = lv["g"].data_ptr()
    pix, interior = lv["pix"].data_ptr(), lv["interior"].data_ptr()
    shifts = [s for (s, _) in lv["taps"]]
    inv = ("scal", 1)
    for sv in reversed(lv["blocks_sv"]):
        blk, bnA, bnB = sv["blk"], sv["bnA"], sv["bnB"]
        # x_out = x_in + BN2(conv2(h)); h = GELU(BN1(conv1(x_in)))
        self._before_write(lv["dzB"].data_ptr())
        self._op(ops, "fvit_bn_bwd", g, 0, Cc, pix, sv["rawB"].data_ptr(), ld, pix, npix, Cc, bnB["mu"].data_ptr(),
                 bnB["rs"].data_ptr(), blk.norm2.weight.data_ptr(), blk.norm2.bias.data_ptr(), L.ACT_NONE, None,
                 bnB["s12"][0].data_ptr(), bnB["s12"][1].data_ptr(), inv, lv["dzB"].data_ptr(), ld, pix,
                 self.G(blk.norm2.weight), self.G(blk.norm2.bias), sv["rs"])
        # h = GELU(BN1(rawA)) was kept by the forward; the pre-activation BN1(rawA) is re-derived inside the dgrad
        # epilogue from rawA and the batch-statistics affine (aux_scale / aux_shift).
        # conv biases sit in front of a batch-statistics BatchNorm: their gradient sum_rows(dz) is identically zero
        # (BN backward removes the per-channel mean), so it is left at the zero the buffer was cleared to.
        _conv_wgrad(self, dz=lv["dzB"].data_ptr(), lddz=ld, x=sv["h16"].data_ptr(), ldx=ld, x_rows=rows, rows=rows, cout=Cc,
                    cin=Cc, shifts=shifts, conv_weight=blk.conv2.weight, name="convB")
        self._bgemm(a=lv["dzB"].data_ptr(), a_rows=rows, lda=ld, b=sv["wT2"].data_ptr(), b_rows=Cc, ldb=9 * _ru(Cc, 64),
                    m=rows, n=Cc, kc=Cc, taps=lv["taps"], act=L.ACT_GELU_BWD, aux=sv["rawA"].data_ptr(), ld_aux=ld,
                    aux_scale=bnA["sc"].data_ptr(), aux_shift=bnA["sh"].data_ptr(),
                    row_map=interior, out_f16=lv["dyA"].data_ptr(), ld_o16=ld, flops=2.0 * npix * Cc * Cc * 9)
        self._before_write(lv["dzA"].data_ptr())
        self._op(ops, "fvit_bn_bwd", lv["dyA"].data_ptr(), 1, ld, pix, sv["rawA"].data_ptr(), ld, pix, npix, Cc,
                 bnA["mu"].data_ptr(), bnA["rs"].data_ptr(), blk.norm1.weight.data_ptr(), blk.norm1.bias.data_ptr(), L.ACT_NONE,
                 None, bnA["s12"][0].data_ptr(), bnA["s12"][1].data_ptr(), inv, lv["dzA"].data_ptr(), ld, pix,
                 self.G(blk.norm1.weight), self.G(blk.norm1.bias), None)
        _conv_wgrad(self, dz=lv["dzA"].data_ptr(), lddz=ld, x=sv["xin"].data_ptr(), ldx=ld, x_rows=rows, rows=rows, cout=Cc,
                    cin=Cc, shifts=shifts, conv_weight=blk.conv1.weight, name="convA")
        # g += conv1 data gradient
        self._bgemm(a=lv["dzA"].data_ptr(), a_rows=rows, lda=ld, b=sv["wT1"].data_ptr(), b_rows=Cc, ldb=9 * _ru(Cc, 64),
                    m=rows, n=Cc, kc=Cc, taps=lv["taps"], row_map=interior, resid=g, ld_resid=Cc, out_f32=g, ld_o32=Cc,
                    flops=2.0 * npix * Cc * Cc * 9)


def _emit_conv_part_bwd(self) -> None:
    ops = self.bwd_ops
    m, B = self.model, self.B
    for idx in range(len(self.conv_lv) - 1, -1, -1):
        lv = self.conv_lv[idx]
        _emit_conv_blocks_bwd(self, lv)
        if idx > 0:
            _emit_downsample_bwd(self, lv["ds"], self.conv_lv[idx - 1], lv)
    # PatchEmbed
    st, lv0 = self.stem_sv, self.conv_lv[0]
    pe = m.patch_embed.conv_down
    nb = self.bufs
    dim, in_dim = lv0["C"], m.cfg["in_dim"]
    H0, W0, pl_rows, ld_in, npix1 = st["H0"], st["W0"], st["pl_rows"], st["ld_in"], st["npix1"]
    inv = ("scal", 1)
    bb, aa, cc = torch.meshgrid(torch.arange(B), torch.arange(H0), torch.arange(W0), indexing="ij")
    q_of_pix = nb.i32("stem.q_of_pix", (bb * (H0 + 1) * (W0 + 1) + (aa + 1) * (W0 + 1) + cc + 1).reshape(-1))
    dY2 = nb.new("stem.dY2", (pl_rows, lv0["ld"]), torch.float16)
    bn2, bn1 = st["bn2"], st["bn1"]
    self._op(ops, "fvit_bn_bwd", lv0["g"].data_ptr(), 0, dim, lv0["pix"].data_ptr(), st["raw2"].data_ptr(), lv0["ld"],
             lv0["pix"].data_ptr(), B * H0 * W0, dim, bn2["mu"].data_ptr(), bn2["rs"].data_ptr(), pe[4].weight.data_ptr(),
             pe[4].bias.data_ptr(), L.ACT_RELU, None, bn2["s12"][0].data_ptr(), bn2["s12"][1].data_ptr(), inv,
             dY2.data_ptr(), lv0["ld"], q_of_pix.data_ptr(), self.G(pe[4].weight), self.G(pe[4].bias), None)
    taps = self._s2_taps(W0)
    shifts = [p * pl_rows + s for (s, p) in taps]
    _conv_wgrad(self, dz=dY2.data_ptr(), lddz=lv0["ld"], x=st["planes"].data_ptr(), ldx=ld_in, x_rows=4 * pl_rows,
                rows=pl_rows, cout=dim, cin=in_dim, shifts=shifts, conv_weight=pe[3].weight, name="stem2")
    co_pad = _ru(dim, 64)
    dplanes = nb.new("stem.dplanes16", (4 * pl_rows, ld_in), torch.float16)
    for p in range(4):
        tl_ = [t for t in range(9) if taps[t][1] == p]
        buf = nb.new(f"stem.wT.{p}", (in_dim, len(tl_) * co_pad), torch.float16)
        t9 = tl_ + [0] * (9 - len(tl_))
        self._op(self.prep_ops, "fvit_pack_conv3x3_taps_f16", pe[3].weight.data_ptr(), buf.data_ptr(), dim, in_dim, co_pad, 1,
                 len(tl_), *t9)
        self._bgemm(a=dY2.data_ptr(), a_rows=pl_rows, lda=lv0["ld"], b=buf.data_ptr(), b_rows=in_dim, ldb=len(tl_) * co_pad,
                    m=pl_rows, n=in_dim, kc=dim, taps=[(-taps[t][0], 0) for t in tl_],
                    out_f16=dplanes.data_ptr() + 2 * p * pl_rows * ld_in, ld_o16=ld_in,
                    flops=2.0 * B * H0 * W0 * in_dim * dim * len(tl_))
    dz1 = nb.new("stem.dz1", (npix1, ld_in), torch.float16)
    self._op(ops, "fvit_bn_bwd", dplanes.data_ptr(), 1, ld_in, st["stem_map"].data_ptr(), st["raw1"].data_ptr(), ld_in,
             st["stem_map"].data_ptr(), npix1, in_dim, bn1["mu"].data_ptr(), bn1["rs"].data_ptr(), pe[1].weight.data_ptr(),
             pe[1].bias.data_ptr(), L.ACT_RELU, None, bn1["s12"][0].data_ptr(), bn1["s12"][1].data_ptr(), inv, dz1.data_ptr(),
             ld_in, None, self.G(pe[1].weight), self.G(pe[1].bias), None)
    self._bgemm(a=dz1.data_ptr(), a_rows=npix1, lda=ld_in, a_mn=True, b=st["col16"].data_ptr(), b_rows=npix1, ldb=self.STEM_LD,
                b_mn=True, m=in_dim, n=27, kc=npix1, split_k=self._split_k(in_dim, 27, npix1), alpha_ptr=inv,
                out_f32=self.G(pe[0].weight), ld_o32=27, flops=2.0 * npix1 * in_dim * 27)


class ConvPartEmitters:
    """TrainPlan mixin: launch-list emitters of PatchEmbed, the conv levels and the Downsamples (the functions above take the
    plan as `self`)."""
    _bn_train = _bn_train
    _bn_finalize = _bn_finalize
    _emit_conv_part_train = _emit_conv_part_train
    _conv_level_train_buffers = _conv_level_train_buffers
    _emit_conv_blocks_train = _emit_conv_blocks_train
    _emit_downsample_train = _emit_downsample_train
    _conv_wgrad = _conv_wgrad
    _emit_downsample_bwd = _emit_downsample_bwd
    _emit_conv_blocks_bwd = _emit_conv_blocks_bwd
    _emit_conv_part_bwd = _emit_conv_part_bwd
