"""Forward(train)/backward emitters of TrainPlan (methods; imported into the class body).

Every function appends launches to plan.ops (train forward) or plan.bwd_ops. Shapes / row maps are the
ones built by Plan (engine.py); `sv` dicts hold the activations saved for the backward pass.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as L
from .engine import _ru


def P(t):
    return None if t is None else t.data_ptr()


def _drop_buf(self, name: str, rate: float, groups: int, group_size: int):
    """row-scale vector (mask / keep per row) of one stochastic-depth site (timm DropPath: one Bernoulli draw
    per leading index — per window for x, per image for carrier tokens and ConvBlocks; fv.py:500,630,652).
    Returns the device pointer or None when the rate is 0."""
    if rate <= 0.0:
        return None
    buf = self.bufs.new(name + ".droppath", (groups * group_size,), torch.float32)
    self.drop_specs.append(dict(name=name, buf=buf, groups=groups, gs=group_size, keep=1.0 - rate))
    return buf.data_ptr()


def _gen_drop_masks(self) -> None:
    forced = getattr(self, "forced_drop_masks", None) or {}
    for d in self.drop_specs:
        if d["name"] in forced:
            m = forced[d["name"]].to(device=d["buf"].device, dtype=torch.float32)
        else:
            m = torch.bernoulli(torch.full((d["groups"],), d["keep"], device=d["buf"].device)) / d["keep"]
        d["buf"].view(d["groups"], d["gs"]).copy_(m.view(-1, 1).expand(d["groups"], d["gs"]))


# ====================================================================================== forward (train)
def _build_forward(self) -> None:
    m, B = self.model, self.B
    cfg = m.cfg
    self.drop_specs: list[dict] = []
    self._block_counter = 0
    self._bn_modules = [mod for mod in m.modules() if isinstance(mod, nn.BatchNorm2d)]
    self.lv: list[dict] = []          # per-level geometry/buffers (+ "sv": saved activations)
    self.ds: list[dict] = []          # per-downsample saved state
    prev = self._emit_conv_part_train()
    Hc, Wc, Cc = prev["H"], prev["W"], prev["C"]
    for i, level in enumerate(m.levels):
        if level.conv:
            continue
        Hc, Wc, Cc = (Hc + 1) // 2, (Wc + 1) // 2, Cc * 2
        tl = self._token_level_buffers(i, level, Cc, Hc, Wc)
        # gradient buffer mirroring xs, plus one guaranteed-zero row at the end
        tl["level_index"] = i
        tl["g"] = self.bufs.new(f"l{i}.g", (tl["xs"].shape[0] + 1, Cc), torch.float32)
        tl["zero_row"] = tl["xs"].shape[0]
        if tl["padded"]:
            self.ops.append(("zero", tl["xs"], "memset"))
        self._emit_downsample_train(i - 1, prev, tl)
        self._emit_token_level_train(i, level, tl)
        tl["kind"] = "tok"
        self.lv.append(tl)
        prev = tl
    self.feat = prev
    self._emit_head_train(prev)


def _posemb_train(self, nm: str, tpe, n_side: int, Cc: int) -> dict:
    """PosEmbMLPSwinv1D in training: same launch as eval plus the saved hidden activations."""
    r = torch.arange(n_side, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(r, r, indexing="ij"))
    grid = (grid - n_side // 2) / (n_side // 2)
    npts = n_side * n_side
    coords = self.bufs.new(nm + ".coords", (npts, 2), torch.float32)
    coords.copy_(grid.flatten(1).t())
    hidden = self.bufs.new(nm + ".hidden", (npts, 512), torch.float32)
    if tuple(tpe.relative_bias.shape) != (1, npts, Cc):
        tpe.relative_bias = torch.zeros(1, npts, Cc, device=self.device)
    self._op(self.prep_ops, "fvit_cpb_mlp_fwd", coords.data_ptr(), npts, tpe.cpb_mlp[0].weight.data_ptr(),
             tpe.cpb_mlp[0].bias.data_ptr(), tpe.cpb_mlp[2].weight.data_ptr(), Cc, tpe.relative_bias.data_ptr(),
             hidden.data_ptr())
    return dict(mod=tpe, coords=coords, hidden=hidden, out=tpe.relative_bias, npts=npts, D=Cc,
                dout=("scr", self._scratch(nm + ".dpe", npts * Cc)))


def _bias_train(self, nm: str, rpb, S: int) -> dict:
    """PosEmbMLPSwinv2D in training: table MLP (hidden saved) + gather/sigmoid into relative_bias."""
    npts = rpb.relative_coords_table.shape[1] * rpb.relative_coords_table.shape[2]
    table = self.bufs.new(nm + ".table", (npts, rpb.num_heads), torch.float32)
    hidden = self.bufs.new(nm + ".hidden", (npts, 512), torch.float32)
    if tuple(rpb.relative_bias.shape) != (1, rpb.num_heads, S, S):
        rpb.relative_bias = torch.zeros(1, rpb.num_heads, S, S, device=self.device)
    self._op(self.prep_ops, "fvit_cpb_mlp_fwd", rpb.relative_coords_table.data_ptr(), npts,
             rpb.cpb_mlp[0].weight.data_ptr(), rpb.cpb_mlp[0].bias.data_ptr(), rpb.cpb_mlp[2].weight.data_ptr(),
             rpb.num_heads, table.data_ptr(), hidden.data_ptr())
    self._op(self.prep_ops, "fvit_attn_bias_fwd", table.data_ptr(), rpb.relative_position_index.data_ptr(),
             rpb.num_heads, S, rpb.window ** 2, rpb.relative_bias.data_ptr())
    return dict(mod=rpb, hidden=hidden, out=rpb.relative_bias, npts=npts, S=S,
                dbias=("scr", self._scratch(nm + ".dbias", rpb.num_heads * S * S)),
                dtable=("scr", self._scratch(nm + ".dtable", npts * rpb.num_heads)))


def _attn_fwd_train(self, nm: str, attn, gamma, rows: int, groups: int, S: int, y16, qkv, ao, bias, stream_buf,
                    u16, rs=None) -> dict:
    """Train-mode copy of Plan._emit_attention: same three launches, un-folded layer scale so that the
    branch output u = proj(...) + b can be saved (u16) for the layer-scale gradient."""
    Cc, h, hd = attn.qkv.in_features, attn.num_heads, attn.head_dim
    hdp = self._head_pad(hd)
    kind = self._attn_kind(S, hdp)
    if kind == "tile" and S > 64 and S % 4 == 0:
        # one window per 128-row tile either way; the key-loop forward also writes the log-sum-exp rows its tensor-core
        # backward needs (fvit_attn_tc_bwd stops at 64-row window slots: without this, 65..128-token windows -- any-res
        # models with 8 x 8 ... 11 x 11 windows -- would train through the SIMT backward)
        kind = "loop"
    use_tc = kind != "simt"
    if not use_tc:
        hdp = hd
    Cp = h * hdp
    qb = attn.qkv.bias
    if hdp == hd:
        wq, ldq = self._pack_linear(nm + ".qkv", attn.qkv)
        qb_ptr = P(qb)
    else:
        ldq = _ru(Cc, 8)
        wq = self.bufs.new(nm + ".qkv.w16", (3 * Cp, ldq), torch.float16)
        self._op(self.prep_ops, "fvit_cast_headpad_f16", attn.qkv.weight.data_ptr(), Cc, wq.data_ptr(), ldq, 3 * Cp,
                 Cc, hd, hdp, 1, 0)
        qb_ptr = None
        if qb is not None:
            qbp = self.bufs.new(nm + ".qkv.bias_pad", (3 * Cp,), torch.float32)
            self._op(self.prep_ops, "fvit_vec_headpad_f32", qb.data_ptr(), qbp.data_ptr(), 3 * Cp, hd, hdp)
            qb_ptr = qbp.data_ptr()
    self._gemm(a=y16.data_ptr(), a_rows=rows, lda=Cc, b=wq.data_ptr(), ldb=ldq, m=rows, n=3 * Cp, kc=Cc,
               col_shift=qb_ptr, out_f16=qkv.data_ptr(), ld_o16=3 * Cp)
    self.op_flops[len(self.ops) - 1] = 2.0 * rows * 3 * Cc * Cc
    scale = float(getattr(attn, "scale", hd ** -0.5))   # qk_scale or head_dim ** -0.5 (fv.py:544)
    lse = None
    if kind == "loop":
        lse = self.bufs.new(nm + ".lse", (rows, h), torch.float32)   # log-sum-exp rows, read by the backward kernel
        self._op(self.ops, "fvit_attn_loop_fwd", qkv.data_ptr(), 3 * Cp, groups, S, h, hdp, bias["out"].data_ptr(),
                 scale, ao.data_ptr(), Cp, lse.data_ptr())
    elif use_tc:
        self._op(self.ops, "fvit_attn_tc_fwd", qkv.data_ptr(), 3 * Cp, groups, S, h, hdp, bias["out"].data_ptr(),
                 scale, ao.data_ptr(), Cp)
    else:
        self._op(self.ops, "fvit_attn_core_fwd", qkv.data_ptr(), 3 * Cp, groups, S, h, hd, bias["out"].data_ptr(),
                 scale, ao.data_ptr(), Cp, None)
    self.op_flops[len(self.ops) - 1] = 4.0 * groups * h * S * S * hd
    lin = attn.proj
    n = lin.weight.shape[0]
    if hdp == hd:
        wp, ldp = self._pack_linear(nm + ".proj", lin)
    else:
        ldp = Cp
        wp = self.bufs.new(nm + ".proj.w16", (n, Cp), torch.float16)
        self._op(self.prep_ops, "fvit_cast_headpad_f16", lin.weight.data_ptr(), Cc, wp.data_ptr(), Cp, n, Cp, hd, hdp,
                 0, 1)
    has_ls = isinstance(gamma, torch.Tensor)
    self._gemm_train_branch(a=ao.data_ptr(), lda=Cp, w=wp.data_ptr(), ldw=ldp, rows=rows, n=n, k=Cp, bias=lin.bias,
                            gamma=gamma if has_ls else None, stream_buf=stream_buf, u16=u16 if has_ls else None, rs=rs)
    self.op_flops[len(self.ops) - 1] = 2.0 * rows * n * Cc
    return dict(rs=rs, attn=attn, wq=wq, ldq=ldq, wp=wp, ldp=ldp, hd=hd, hdp=hdp, Cp=Cp, use_tc=use_tc, kind=kind, lse=lse,
                scale=scale, S=S,
                groups=groups, rows=rows, qkv=qkv, ao=ao, y16=y16, bias=bias, u16=u16 if has_ls else None)


def _gemm_train_branch(self, *, a, lda, w, ldw, rows, n, k, bias, gamma, stream_buf, u16, rs=None) -> None:
    """x += droppath_row * gamma * (a @ W^T + b), saving u = a @ W^T + b (fp16) when gamma is a parameter."""
    g = L.GemmArgs()
    g.a, g.a_rows, g.lda, g.a_planes = a, rows, lda, 1
    g.b, g.b_rows, g.ldb = w, n, ldw
    g.m, g.n, g.kc, g.ntaps = rows, n, k, 1
    g.split_k, g.alpha = 1, 1.0
    g.col_shift = P(bias)
    if gamma is not None:
        g.col_scale2 = gamma.data_ptr()
        g.out_pre16, g.ld_out_pre16 = u16.data_ptr(), n
    g.row_scale = rs
    g.resid, g.ld_resid, g.out_f32, g.ld_out_f32 = stream_buf, n, stream_buf, n
    self._gemm_keep.append(g)
    import ctypes as C
    self.ops.append((self.lib.fvit_gemm, (C.byref(g),), "fvit_gemm"))
    self.op_flops[len(self.ops) - 1] = 2.0 * rows * n * k


def _emit_token_level_train(self, i: int, level, tl: dict) -> None:
    B, Cc, S, ncw, ws = self.B, tl["C"], tl["S"], tl["ncw"], tl["ws"]
    nW, n_ct = tl["nW"], tl["n_ct"]
    nb = self.bufs
    xs_ptr = tl["xs"].data_ptr()
    has_ct = ncw > 0
    hid = int(Cc * self.model.cfg["mlp_ratio"])
    heads = level.blocks[0].attn.num_heads
    Cp = heads * self._head_pad(Cc // heads)
    rows = nW * S
    tl["blocks_sv"] = []
    if level.do_gt and has_ct:
        tk = level.global_tokenizer
        (kh, sh_, oh), (kw, sw_, ow) = tk.pool
        self._op(self.ops, "fvit_token_init_fwd", xs_ptr, Cc, tl["pix_map"].data_ptr(), B, tl["Hp"], tl["Wp"], Cc,
                 tk.pos_embed.weight.data_ptr(), tk.pos_embed.bias.data_ptr(), kh, kw, sh_, sw_, oh, ow,
                 tl["ct_rows"].data_ptr(), xs_ptr, Cc)

    def new16(name, r, c):
        return nb.new(name, (r, c), torch.float16)

    def new32(name, r):
        return nb.new(name, (r,), torch.float32)

    for j, blk in enumerate(level.blocks):
        nm = f"l{i}.b{j}"
        key = f"levels.{i}.blocks.{j}"
        rate = self.model.drop_path_rates[self._block_counter]
        self._block_counter += 1
        sv: dict = dict(blk=blk)
        sv["rs_attn"] = self._drop_buf(key + ".attn", rate, nW, S)
        sv["rs_mlp"] = self._drop_buf(key + ".mlp", rate, nW, S)
        if has_ct:
            sv["rs_hat_attn"] = self._drop_buf(key + ".hat_attn", rate, B, n_ct)
            sv["rs_hat_mlp"] = self._drop_buf(key + ".hat_mlp", rate, B, n_ct)
        sv["pe"] = self._posemb_train(nm + ".pe", blk.pos_embed, ws, Cc)
        if has_ct:
            ctr_ptr = xs_ptr + tl["ctr0"] * Cc * 4
            rows_c = B * n_ct
            sv["hat_pe"] = None
            if hasattr(blk, "hat_pos_embed"):
                sv["hat_pe"] = self._posemb_train(nm + ".hat_pe", blk.hat_pos_embed, int(n_ct ** 0.5), Cc)
            c = dict(y1=new16(nm + ".c.y1", rows_c, Cc), xh1=new16(nm + ".c.xh1", rows_c, Cc), rs1=new32(nm + ".c.rs1", rows_c),
                     mu1=new32(nm + ".c.mu1", rows_c), y2=new16(nm + ".c.y2", rows_c, Cc), xh2=new16(nm + ".c.xh2", rows_c, Cc),
                     rs2=new32(nm + ".c.rs2", rows_c), mu2=new32(nm + ".c.mu2", rows_c),
                     qkv=new16(nm + ".c.qkv", rows_c, 3 * Cp), ao=new16(nm + ".c.ao", rows_c, Cp),
                     p=new16(nm + ".c.p", rows_c, hid), h=new16(nm + ".c.h", rows_c, hid),
                     u1=new16(nm + ".c.u1", rows_c, Cc), u2=new16(nm + ".c.u2", rows_c, Cc))
            sv["c"] = c
            self._op(self.ops, "fvit_ln_fwd", xs_ptr, Cc, tl["ct_gather"].data_ptr(), rows_c, Cc,
                     sv["hat_pe"]["out"].data_ptr() if sv["hat_pe"] else None, n_ct, 0, ctr_ptr, Cc,
                     blk.hat_norm1.weight.data_ptr(), blk.hat_norm1.bias.data_ptr(), float(blk.hat_norm1.eps),
                     c["y1"].data_ptr(), Cc, None, c["mu1"].data_ptr(), c["rs1"].data_ptr(), c["xh1"].data_ptr(), Cc)
            sv["c_bias"] = self._bias_train(nm + ".hat_bias", blk.hat_attn.pos_emb_funct, n_ct)
            sv["c_attn"] = self._attn_fwd_train(nm + ".hat_attn", blk.hat_attn, blk.gamma1, rows_c, B, n_ct, c["y1"],
                                                c["qkv"], c["ao"], sv["c_bias"], ctr_ptr, c["u1"], sv["rs_hat_attn"])
            self._op(self.ops, "fvit_ln_fwd", ctr_ptr, Cc, None, rows_c, Cc, None, 1, 0, None, 0,
                     blk.hat_norm2.weight.data_ptr(), blk.hat_norm2.bias.data_ptr(), float(blk.hat_norm2.eps),
                     c["y2"].data_ptr(), Cc, None, c["mu2"].data_ptr(), c["rs2"].data_ptr(), c["xh2"].data_ptr(), Cc)
            sv["c_mlp"] = self._mlp_fwd_train(nm + ".hat_mlp", blk.hat_mlp, blk.gamma2, rows_c, c["y2"], c["p"], c["h"],
                                              ctr_ptr, c["u2"], sv["rs_hat_mlp"])
        w = dict(y1=new16(nm + ".w.y1", rows, Cc), xh1=new16(nm + ".w.xh1", rows, Cc), rs1=new32(nm + ".w.rs1", rows),
                 mu1=new32(nm + ".w.mu1", rows), y2=new16(nm + ".w.y2", rows, Cc), xh2=new16(nm + ".w.xh2", rows, Cc),
                 rs2=new32(nm + ".w.rs2", rows), mu2=new32(nm + ".w.mu2", rows), qkv=new16(nm + ".w.qkv", rows, 3 * Cp),
                 ao=new16(nm + ".w.ao", rows, Cp), p=new16(nm + ".w.p", rows, hid), h=new16(nm + ".w.h", rows, hid),
                 u1=new16(nm + ".w.u1", rows, Cc), u2=new16(nm + ".w.u2", rows, Cc))
        sv["w"] = w
        self._op(self.ops, "fvit_ln_fwd", xs_ptr, Cc, tl["norm1_gather"].data_ptr() if has_ct else None, rows, Cc,
                 sv["pe"]["out"].data_ptr(), S, ncw, xs_ptr, Cc, blk.norm1.weight.data_ptr(), blk.norm1.bias.data_ptr(),
                 float(blk.norm1.eps), w["y1"].data_ptr(), Cc, None, w["mu1"].data_ptr(), w["rs1"].data_ptr(),
                 w["xh1"].data_ptr(), Cc)
        sv["w_bias"] = self._bias_train(nm + ".bias", blk.attn.pos_emb_funct, S)
        sv["w_attn"] = self._attn_fwd_train(nm + ".attn", blk.attn, blk.gamma3, rows, nW, S, w["y1"], w["qkv"], w["ao"],
                                            sv["w_bias"], xs_ptr, w["u1"], sv["rs_attn"])
        self._op(self.ops, "fvit_ln_fwd", xs_ptr, Cc, None, rows, Cc, None, 1, 0, None, 0, blk.norm2.weight.data_ptr(),
                 blk.norm2.bias.data_ptr(), float(blk.norm2.eps), w["y2"].data_ptr(), Cc, None, w["mu2"].data_ptr(),
                 w["rs2"].data_ptr(), w["xh2"].data_ptr(), Cc)
        sv["w_mlp"] = self._mlp_fwd_train(nm + ".mlp", blk.mlp, blk.gamma4, rows, w["y2"], w["p"], w["h"], xs_ptr, w["u2"],
                                          sv["rs_mlp"])
        if has_ct and blk.last and blk.do_propagation:
            g1 = blk.gamma1.data_ptr() if isinstance(blk.gamma1, torch.Tensor) else None
            self._op(self.ops, "fvit_propagate_fwd", xs_ptr, Cc, tl["prop_src"].data_ptr(), rows, Cc, g1)
            sv["prop"] = True
        tl["blocks_sv"].append(sv)
    # level-wide backward scratch
    tl["dz"] = nb.new(f"l{i}.dz", (max(rows, 1), Cc), torch.float16)
    tl["dy"] = nb.new(f"l{i}.dy", (max(rows, 1), Cc), torch.float16)
    tl["dp"] = nb.new(f"l{i}.dp", (max(rows, 1), hid), torch.float16)
    tl["dao"] = nb.new(f"l{i}.dao", (max(rows, 1), Cp), torch.float16)
    tl["dqkv"] = nb.new(f"l{i}.dqkv", (max(rows, 1), 3 * Cp), torch.float16)


def _mlp_fwd_train(self, nm: str, mlp, gamma, rows: int, y16, p16, h16, stream_buf, u16, rs=None) -> dict:
    n1, k1 = mlp.fc1.weight.shape
    w1, ld1 = self._pack_linear(nm + ".fc1", mlp.fc1)
    # fc1: gelu'(pre-activation) saved through out_pre16 (one erf evaluation yields GELU and its derivative, so the
    # fc2 data-gradient epilogue is a plain multiply: FVIT_ACT_MUL_AUX), GELU output is the fc2 operand
    g = L.GemmArgs()
    g.a, g.a_rows, g.lda, g.a_planes = y16.data_ptr(), rows, k1, 1
    g.b, g.b_rows, g.ldb = w1.data_ptr(), n1, ld1
    g.m, g.n, g.kc, g.ntaps, g.split_k, g.alpha = rows, n1, k1, 1, 1, 1.0
    g.col_shift, g.act = mlp.fc1.bias.data_ptr(), L.ACT_GELU
    g.out_f16, g.ld_out_f16 = h16.data_ptr(), n1
    g.out_pre16, g.ld_out_pre16 = p16.data_ptr(), n1
    g.pre_is_grad = 1
    self._gemm_keep.append(g)
    import ctypes as C
    self.ops.append((self.lib.fvit_gemm, (C.byref(g),), "fvit_gemm"))
    self.op_flops[len(self.ops) - 1] = 2.0 * rows * n1 * k1
    n2, k2 = mlp.fc2.weight.shape
    w2, ld2 = self._pack_linear(nm + ".fc2", mlp.fc2)
    has_ls = isinstance(gamma, torch.Tensor)
    self._gemm_train_branch(a=h16.data_ptr(), lda=k2, w=w2.data_ptr(), ldw=ld2, rows=rows, n=n2, k=k2,
                            bias=mlp.fc2.bias, gamma=gamma if has_ls else None, stream_buf=stream_buf,
                            u16=u16 if has_ls else None, rs=rs)
    return dict(rs=rs, mlp=mlp, w1=w1, ld1=ld1, w2=w2, ld2=ld2, rows=rows, y16=y16, p16=p16, h16=h16,
                u16=u16 if has_ls else None)


def _emit_head_train(self, prev: dict) -> None:
    """BatchNorm2d with batch statistics folded into the pool (fv.py:953-958 under .train())."""
    m, B, nb = self.model, self.B, self.bufs
    nf = m.num_features
    T = prev["H"] * prev["W"]
    st = nb.new("norm.stats", (2, nf), torch.float32)
    self.fwd_zero.append(st)
    sN = nb.new("norm.scale", (nf,), torch.float32)
    tN = nb.new("norm.shift", (nf,), torch.float32)
    mu = nb.new("norm.mean", (nf,), torch.float32)
    rs = nb.new("norm.rstd", (nf,), torch.float32)
    self._op(self.ops, "fvit_colstats_f32", prev["xs"].data_ptr(), nf, prev["crop_map"].data_ptr(), B * T, nf,
             st[0].data_ptr(), st[1].data_ptr())
    self._op(self.ops, "fvit_bn_finalize", st[0].data_ptr(), st[1].data_ptr(), float(B * T), m.norm.weight.data_ptr(),
             m.norm.bias.data_ptr(), float(m.norm.eps), float(m.norm.momentum), m.norm.running_mean.data_ptr(),
             m.norm.running_var.data_ptr(), None, sN.data_ptr(), tN.data_ptr(), mu.data_ptr(), rs.data_ptr(), nf)
    pooled = nb.new("head.pooled", (B, _ru(nf, 8)), torch.float16)
    self._op(self.ops, "fvit_pool_affine_fwd", prev["xs"].data_ptr(), nf, prev["crop_map"].data_ptr(), B, T, nf,
             sN.data_ptr(), tN.data_ptr(), pooled.data_ptr(), pooled.stride(0))
    hw16, ldh = self._pack_linear("head", m.head)
    self.logits = nb.new("logits", (B, m.num_classes), torch.float32)
    self._gemm(a=pooled.data_ptr(), a_rows=B, lda=pooled.stride(0), b=hw16.data_ptr(), ldb=ldh, m=B, n=m.num_classes,
               kc=nf, col_shift=m.head.bias.data_ptr(), out_f32=self.logits.data_ptr(), ld_o32=m.num_classes)
    self.head_sv = dict(pooled=pooled, hw16=hw16, ldh=ldh, mu=mu, rs=rs, T=T, nf=nf)


# ====================================================================================== backward
def _build_backward(self) -> None:
    m = self.model
    # gradient buckets of the data-parallel all-reduce, in completion order. The flat buffer follows
    # .parameters() order (patch_embed, levels.0 .. levels.n, norm, head); the backward pass finishes
    # [levels.n, norm, head] first, then one transformer level at a time, then the conv part.
    # Transformer levels are cut further: a bucket point after every group of blocks whose gradients add up to
    # FVIT_BUCKET_MB (default 100 MB; fv4: every 2 level-2 blocks / every level-3 block), so that the reduction of a
    # level overlaps that level's own backward and no single burst lands on the few large conv weight-gradient launches
    # at the end. The rest of a level (downsample, tokenizer) goes with the level-end point.
    import os
    first = {}
    for name, p in m.named_parameters():
        if name.startswith("levels."):
            first.setdefault(int(name.split(".")[1]), self._goff[id(p)])
    total = self.gflat.numel()
    group_floats = int(float(os.environ.get("FVIT_BUCKET_MB", "100")) * 1e6 / 4)
    self.grad_buckets = []

    def point(lo: int, hi: int) -> None:
        if hi > lo:
            self.grad_buckets.append((lo, hi))
            self.bwd_ops.append(("bucket", (lo, hi), "grad_bucket"))

    self._emit_head_bwd(self.feat)
    toks = [lv for lv in self.lv]
    n_conv = len(m.levels) - len(toks)
    hi = total
    for idx in range(len(toks) - 1, -1, -1):
        tl = toks[idx]
        lo = first.get(n_conv + idx, hi)
        # offsets of the level's blocks in the flat buffer (module order = offset order, contiguous)
        spans = []
        for blk in m.levels[tl["level_index"]].blocks:
            offs = [(self._goff[id(q)], q.numel()) for q in blk.parameters()]
            spans.append((min(o for o, _ in offs), max(o + _ru(n, 64) for o, n in offs)))
        contiguous = all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans and lo <= spans[0][0] and spans[-1][1] <= hi
        state = dict(top=spans[-1][1] if spans else hi)   # blocks above `top` have been handed to a bucket already

        def after_block(k: int, spans=spans, state=state) -> None:
            if group_floats > 0 and contiguous and k > 0 and state["top"] - spans[k][0] >= group_floats:
                point(spans[k][0], state["top"])
                state["top"] = spans[k][0]
        self._emit_token_level_bwd(tl, after_block)
        if contiguous and state["top"] < spans[-1][1]:   # some block groups are out already: the rest in two pieces
            point(spans[-1][1], hi)
            point(lo, state["top"])
        else:
            point(lo, hi)
        hi = lo
        src = toks[idx - 1] if idx > 0 else self.conv_out
        self._emit_downsample_bwd(tl["ds"], src, tl)
    self._emit_conv_part_bwd()
    point(0, hi)


def _emit_head_bwd(self, feat: dict) -> None:
    m, B, nb = self.model, self.B, self.bufs
    hs = self.head_sv
    nf, ncls, T = hs["nf"], m.num_classes, hs["T"]
    ldl = _ru(ncls, 8)
    ncp = _ru(ncls, 4)   # the loss gradient is kept in rows padded to 16 bytes (any num_classes; padding stays zero)
    self._dlogits = nb.new("grad.dlogits", (B, ncp), torch.float32)
    self._ncls = ncls
    dl16 = nb.new("grad.dl16", (B, ldl), torch.float16)
    dpool = nb.new("grad.dpool", (B, nf), torch.float32)
    inv = ("scal", 1)
    ops = self.bwd_ops
    self._op(ops, "fvit_cast_scale_f16", self._dlogits.data_ptr(), ncp, None, B, ncp, None, ("scal", 0),
             dl16.data_ptr(), ldl, None)
    self._op(ops, "fvit_colsum", self._dlogits.data_ptr(), 0, ncp, None, None, 0, B, ncls, None, None,
             self.G(m.head.bias), None)
    # dW_head = dl^T pooled ; dpooled = dl W_head
    self._bgemm(a=dl16.data_ptr(), a_rows=B, lda=ldl, a_mn=True, b=hs["pooled"].data_ptr(), b_rows=B,
                ldb=hs["pooled"].stride(0), b_mn=True, m=ncls, n=nf, kc=B, alpha_ptr=inv, out_f32=self.G(m.head.weight),
                ld_o32=nf)
    self._bgemm(a=dl16.data_ptr(), a_rows=B, lda=ldl, b=hs["hw16"].data_ptr(), b_rows=ncls, ldb=hs["ldh"], b_mn=True,
                m=B, n=nf, kc=ncls, out_f32=dpool.data_ptr(), ld_o32=nf)
    s12 = nb.new("grad.head_s12", (2, nf), torch.float32)
    self._op(ops, "fvit_pool_bn_bwd", feat["xs"].data_ptr(), nf, feat["crop_map"].data_ptr(), B, T, nf,
             hs["mu"].data_ptr(), hs["rs"].data_ptr(), m.norm.weight.data_ptr(), dpool.data_ptr(), nf,
             s12[0].data_ptr(), s12[1].data_ptr(), inv, feat["g"].data_ptr(), nf, self.G(m.norm.weight),
             self.G(m.norm.bias))


def _emit_attn_core_bwd(self, at: dict, dao, dqkv, groups: int, S: int) -> None:
    """backward of the attention core: the tcgen05 tile kernel (S <= 64), the tcgen05 key-loop kernel (S > 128: needs the
    forward's output and log-sum-exp; per-pair partial dQ reduced into an fp32 scratch matrix -- any-res S = 148, the
    14 x 14 ... 48 x 48 windows of the 21k models) or the generic SIMT kernel. (fvit_attn_loop_bwd, the two-tile variant
    that keeps all of dQ in TMEM, measured 25 % slower on S = 148 / 196 than the eight-softmax-warp kernel: r02t.)"""
    attn = at["attn"]
    h, hd, hdp, Cp = attn.num_heads, at["hd"], at["hdp"], at["Cp"]
    ops = self.bwd_ops
    self._before_write(dqkv.data_ptr())
    if at.get("kind") == "loop":
        # one scratch matrix per plan, sized for the largest attention of the model (the blocks' backward passes run
        # one after the other on the main stream; the kernel leaves nothing in it)
        need = groups * S * Cp
        scr = getattr(self, "_dq_scratch", None)
        if scr is None or scr.numel() < need:
            scr = self.bufs.new(f"grad.attn_dq32.{need}", (need,), torch.float32)   # (a smaller one stays alive: its
            self._dq_scratch = scr                                                  # launches keep their pointer)
        self._op(ops, "fvit_attn_loop_bwd_long", at["qkv"].data_ptr(), 3 * Cp, dao.data_ptr(), Cp, at["ao"].data_ptr(), Cp,
                 at["lse"].data_ptr(), groups, S, h, hdp, at["bias"]["out"].data_ptr(), at["scale"], dqkv.data_ptr(), 3 * Cp,
                 at["bias"]["dbias"], scr.data_ptr(), Cp)
    else:
        self._op(ops, _attn_bwd_entry(S, hdp, at["use_tc"]), at["qkv"].data_ptr(), 3 * Cp, dao.data_ptr(), Cp, groups, S, h, hd,
                 hdp, at["bias"]["out"].data_ptr(), at["scale"], dqkv.data_ptr(), 3 * Cp, at["bias"]["dbias"])
    self.bwd_flops[len(ops) - 1] = 10.0 * groups * h * S * S * hd


def _attn_bwd_entry(S: int, hdp: int, use_tc: bool) -> str:
    """tensor-core attention backward when the tile fits (S <= 64, head slices of 32/64, shared memory budget)"""
    if use_tc and S <= 64 and hdp in (32, 64):
        ss = (S * S * 4 + 15) // 16 * 16
        if 1024 + 2 * 4 * 128 * hdp * 2 + 2 * 128 * 128 * 2 + 2 * ss + 128 <= 227 * 1024:
            return "fvit_attn_tc_bwd"
    return "fvit_attn_core_bwd"


def _posemb_bwd(self, pe: dict) -> None:
    mod = pe["mod"]
    n0 = len(self.bwd_ops)
    self._op(self.bwd_ops, "fvit_cpb_mlp_bwd", pe["coords"].data_ptr(), pe["npts"], mod.cpb_mlp[2].weight.data_ptr(),
             pe["hidden"].data_ptr(), pe["dout"], pe["D"], None, self.G(mod.cpb_mlp[0].weight),
             self.G(mod.cpb_mlp[0].bias), self.G(mod.cpb_mlp[2].weight))
    self._bwd_side.update(range(n0, len(self.bwd_ops)))


def _bias_bwd(self, bs: dict) -> None:
    """Gradient of the relative-position bias MLP. Its inputs (this block's dbias / dtable scratch slices) are written
    once per backward pass and its outputs are parameter gradients nobody downstream reads: a leaf chain, run as a
    side branch of the launch graph (TrainPlan._bwd_side) so that these small latency-bound launches overlap the GEMMs."""
    rpb = bs["mod"]
    n0 = len(self.bwd_ops)
    self._op(self.bwd_ops, "fvit_attn_bias_bwd", bs["dbias"], bs["out"].data_ptr(), rpb.relative_position_index.data_ptr(),
             rpb.num_heads, bs["S"], rpb.window ** 2, ("scal", 1), bs["dtable"])
    self._op(self.bwd_ops, "fvit_cpb_mlp_bwd", rpb.relative_coords_table.data_ptr(), bs["npts"],
             rpb.cpb_mlp[2].weight.data_ptr(), bs["hidden"].data_ptr(), bs["dtable"], rpb.num_heads, None,
             self.G(rpb.cpb_mlp[0].weight), self.G(rpb.cpb_mlp[0].bias), self.G(rpb.cpb_mlp[2].weight))
    self._bwd_side.update(range(n0, len(self.bwd_ops)))


def _mlp_bwd(self, tl: dict, ms: dict, gamma, g_ptr: int, rows: int, ln: nn.LayerNorm, xh, rs, in_map=None) -> None:
    """Backward of x += gamma * fc2(gelu(fc1(LN(x)))) on gradient rows g_ptr[0:rows]."""
    ops = self.bwd_ops
    mlp = ms["mlp"]
    Cc, hid = mlp.fc1.weight.shape[1], mlp.fc1.weight.shape[0]
    br = self._branch(gamma)
    dz, dp, dy = tl["dz"], tl["dp"], tl["dy"]
    fused = Cc % 8 == 0 and mlp.fc2.bias is not None
    self._before_write(dz.data_ptr())
    if fused:   # operand cast + fc2 bias gradient + layer-scale gradient in one pass over g
        has_g = br["gamma"] is not None
        self._op(ops, "fvit_branch_grad", g_ptr, Cc, rows, Cc, P(br["gamma"]), br["s"], ms["rs"], dz.data_ptr(), Cc,
                 br["w_alpha"], self.G(mlp.fc2.bias), ms["u16"].data_ptr() if has_g else None, Cc, ("scal", 1),
                 self.G(gamma) if has_g else None)
    else:
        self._op(ops, "fvit_cast_scale_f16", g_ptr, Cc, None, rows, Cc, P(br["gamma"]), br["s"], dz.data_ptr(), Cc, ms["rs"])
        if br["gamma"] is not None:
            self._op(ops, "fvit_colsum", g_ptr, 0, Cc, None, ms["u16"].data_ptr(), Cc, rows, Cc, None, ("scal", 1),
                     self.G(gamma), ms["rs"])
    # fc2: dW2, db2, dp = (dz W2) o gelu'(p)   (p16 holds gelu'(p), saved by the forward epilogue)
    self._linear_bwd(lin=mlp.fc2, w16=ms["w2"].data_ptr(), ldw=ms["ld2"], x16=ms["h16"].data_ptr(), ldx=hid,
                     dz16=dz.data_ptr(), lddz=Cc, rows=rows, n_out=Cc, k_in=hid, br=br, dx16=dp.data_ptr(), lddx=hid,
                     dx_act=L.ACT_MUL_AUX, dx_aux=ms["p16"].data_ptr(), ld_aux=hid, bias_done=fused,
                     dx_colsum=(self.G(mlp.fc1.bias), ("scal", 1)) if mlp.fc1.bias is not None else None)
    one = dict(gamma=None, s=None, inv_s=("scal", 2), w_alpha=("scal", 1))
    # (fc1's bias gradient = column sums of dp, taken in the epilogue of the GEMM that produced dp)
    self._linear_bwd(lin=mlp.fc1, w16=ms["w1"].data_ptr(), ldw=ms["ld1"], x16=ms["y16"].data_ptr(), ldx=Cc,
                     dz16=dp.data_ptr(), lddz=hid, rows=rows, n_out=hid, k_in=Cc, br=one, dx16=dy.data_ptr(), lddx=Cc,
                     bias_done=mlp.fc1.bias is not None)
    self._op(ops, "fvit_ln_bwd", dy.data_ptr(), Cc, None, xh.data_ptr(), Cc, rs.data_ptr(), ln.weight.data_ptr(), rows, Cc,
             g_ptr, Cc, in_map, 1, 0, ("scal", 1), self.G(ln.weight), self.G(ln.bias))


def _attn_bwd(self, tl: dict, at: dict, gamma, g_ptr: int, ln: nn.LayerNorm, xh, rs, in_map, g_base: int,
              clear_moved: bool) -> None:
    """Backward of x += gamma * proj(attn(qkv(LN(x)))); g_base = pointer the in_map row indices refer to."""
    ops = self.bwd_ops
    attn = at["attn"]
    Cc, h, hd, hdp, Cp = attn.qkv.in_features, attn.num_heads, at["hd"], at["hdp"], at["Cp"]
    rows, S, groups = at["rows"], at["S"], at["groups"]
    br = self._branch(gamma)
    dz, dy, dao, dqkv = tl["dz"], tl["dy"], tl["dao"], tl["dqkv"]
    padded = hdp != hd
    fused = Cc % 8 == 0 and attn.proj.bias is not None
    self._before_write(dz.data_ptr())
    if fused:   # operand cast + proj bias gradient + layer-scale gradient in one pass over g
        has_g = br["gamma"] is not None
        self._op(ops, "fvit_branch_grad", g_ptr, Cc, rows, Cc, P(br["gamma"]), br["s"], at["rs"], dz.data_ptr(), Cc,
                 br["w_alpha"], self.G(attn.proj.bias), at["u16"].data_ptr() if has_g else None, Cc, ("scal", 1),
                 self.G(gamma) if has_g else None)
    else:
        self._op(ops, "fvit_cast_scale_f16", g_ptr, Cc, None, rows, Cc, P(br["gamma"]), br["s"], dz.data_ptr(), Cc, at["rs"])
        if br["gamma"] is not None:
            self._op(ops, "fvit_colsum", g_ptr, 0, Cc, None, at["u16"].data_ptr(), Cc, rows, Cc, None, ("scal", 1),
                     self.G(gamma), at["rs"])
    # proj
    if padded:
        gWp = ("scr", self._scratch("dWproj_pad", Cc * Cp))
    else:
        gWp = self.G(attn.proj.weight)
    self._linear_bwd(lin=attn.proj, w16=at["wp"].data_ptr(), ldw=at["ldp"], x16=at["ao"].data_ptr(), ldx=Cp,
                     dz16=dz.data_ptr(), lddz=Cc, rows=rows, n_out=Cc, k_in=Cp, br=br, gW=gWp, gW_ld=Cp,
                     dx16=dao.data_ptr(), lddx=Cp, flops_k=Cc, bias_done=fused)
    if padded:
        self._op(ops, "fvit_unpad_heads_f32", gWp, Cp, self.G(attn.proj.weight), Cc, Cc, Cp, hd, hdp, 0, 1, None)
        self._bwd_side.add(len(ops) - 1)   # fresh scratch in, parameter gradient out: leaf launch
    # attention core
    _emit_attn_core_bwd(self, at, dao, dqkv, groups, S)
    _bias_bwd(self, at["bias"])
    # qkv
    one = dict(gamma=None, s=None, inv_s=("scal", 2), w_alpha=("scal", 1))
    gWq, gbq = self.G(attn.qkv.weight), (self.G(attn.qkv.bias) if attn.qkv.bias is not None else None)
    if padded:   # dW rows are head-padded: the GEMM epilogue scatters them through a row map, only the bias is un-padded
        gbq = ("scr", self._scratch("dbqkv_pad", 3 * Cp))
    self._linear_bwd(lin=attn.qkv, w16=at["wq"].data_ptr(), ldw=at["ldq"], x16=at["y16"].data_ptr(), ldx=Cc,
                     dz16=dqkv.data_ptr(), lddz=3 * Cp, rows=rows, n_out=3 * Cp, k_in=Cc, br=one, gW=gWq, gW_ld=Cc,
                     gW_row_map=_unpad_row_map(self, 3 * h, hd, hdp) if padded else None,
                     bias_to=gbq, dx16=dy.data_ptr(), lddx=Cc, flops_k=Cc * (3 * Cc) / (3 * Cp))
    if padded:
        if attn.qkv.bias is not None:
            self._op(ops, "fvit_unpad_heads_f32", gbq, 1, self.G(attn.qkv.bias), 1, 3 * Cp, 1, hd, hdp, 1, 0, None)
            self._bwd_side.add(len(ops) - 1)   # fresh scratch in, parameter gradient out: leaf launch
    # LayerNorm (with the gather routing of the forward)
    self._op(ops, "fvit_ln_bwd", dy.data_ptr(), Cc, None, xh.data_ptr(), Cc, rs.data_ptr(), ln.weight.data_ptr(), rows, Cc,
             g_ptr, Cc, in_map, 1, 1 if clear_moved else 0, ("scal", 1), self.G(ln.weight), self.G(ln.bias))


def _unpad_row_map(self, heads3: int, hd: int, hdp: int) -> int:
    """device int32 map: row r of a head-padded [heads3 * hdp, .] matrix -> row of the un-padded [heads3 * hd, .]
    parameter layout, -1 for padding rows (lets the qkv weight-gradient GEMM scatter straight into the flat gradient
    buffer instead of writing a padded scratch matrix that a second pass un-pads)."""
    cache = self.__dict__.setdefault("_unpad_maps", {})
    key = (heads3, hd, hdp)
    if key not in cache:
        r = torch.arange(heads3 * hdp)
        m = torch.where(r % hdp < hd, (r // hdp) * hd + r % hdp, torch.full_like(r, -1))
        cache[key] = self.bufs.i32(f"unpad_rows.{heads3}.{hd}.{hdp}", m)
    return cache[key].data_ptr()


def _emit_token_level_bwd(self, tl: dict, after_block=None) -> None:
    B, Cc, S, ncw, ws = self.B, tl["C"], tl["S"], tl["ncw"], tl["ws"]
    nW, n_ct = tl["nW"], tl["n_ct"]
    ops = self.bwd_ops
    g_ptr = tl["g"].data_ptr()
    has_ct = ncw > 0
    rows = nW * S
    for sv in reversed(tl["blocks_sv"]):
        blk = sv["blk"]
        w = sv["w"]
        if sv.get("prop"):
            g1 = blk.gamma1 if isinstance(blk.gamma1, torch.Tensor) else None
            self._op(ops, "fvit_propagate_bwd", g_ptr, Cc, tl["xs"].data_ptr(), Cc, tl["prop_src"].data_ptr(), rows, Cc,
                     P(g1), ("scal", 1), self.G(g1) if g1 is not None else None)
        # window branch (reverse order): MLP, attention + norm1 (routes carrier-slot gradients to the raster buffer)
        _mlp_bwd(self, tl, sv["w_mlp"], blk.gamma4, g_ptr, rows, blk.norm2, w["xh2"], w["rs2"])
        _attn_bwd(self, tl, sv["w_attn"], blk.gamma3, g_ptr, blk.norm1, w["xh1"], w["rs1"],
                  tl["norm1_gather"].data_ptr() if has_ct else None, g_ptr, True)
        # positional embedding of the window tokens: sum over windows of the gradient at (x + pe)
        self._op(ops, "fvit_group_sum", g_ptr, Cc, nW, S, ncw, Cc, ("scal", 1), sv["pe"]["dout"])
        _posemb_bwd(self, sv["pe"])
        if has_ct:
            c = sv["c"]
            gc_ptr = g_ptr + tl["ctr0"] * Cc * 4
            rows_c = B * n_ct
            _mlp_bwd(self, tl, sv["c_mlp"], blk.gamma2, gc_ptr, rows_c, blk.hat_norm2, c["xh2"], c["rs2"])
            # carrier attention; its LayerNorm gathered from the xs carrier rows (ct_dewindow): the
            # gradient goes back there (row indices are relative to the level's g buffer)
            _attn_bwd_carrier(self, tl, sv, blk, gc_ptr, g_ptr)
        if after_block is not None:
            after_block(tl["blocks_sv"].index(sv))   # this block's parameter gradients are final
    level = self.model.levels[tl["level_index"]]
    if level.do_gt and has_ct:
        tk = level.global_tokenizer
        (kh, sh_, oh), (kw, sw_, ow) = tk.pool
        self._op(ops, "fvit_token_init_bwd", g_ptr, Cc, tl["x0_16"].data_ptr(), Cc, tl["pix_map"].data_ptr(),
                 tl["ct_rows"].data_ptr(), B, tl["Hp"], tl["Wp"], Cc, tk.pos_embed.weight.data_ptr(), kh, kw, sh_, sw_, oh, ow,
                 ("scal", 1), g_ptr, Cc, self.G(tk.pos_embed.weight), self.G(tk.pos_embed.bias))


def _attn_bwd_carrier(self, tl: dict, sv: dict, blk, gc_ptr: int, g_ptr: int) -> None:
    """hat_attn + hat_norm1 backward. ln_bwd works on row indices of one buffer, so the raster rows are
    addressed as rows of the level buffer (offset ctr0) and in_map = ct_gather (xs carrier rows)."""
    ops = self.bwd_ops
    B, Cc, n_ct = self.B, tl["C"], tl["n_ct"]
    c = sv["c"]
    at = sv["c_attn"]
    rows_c = B * n_ct
    # everything except the final LayerNorm backward is identical to the window case
    attn = at["attn"]
    h, hd, hdp, Cp = attn.num_heads, at["hd"], at["hdp"], at["Cp"]
    br = self._branch(blk.gamma1)
    dz, dy, dao, dqkv = tl["dz"], tl["dy"], tl["dao"], tl["dqkv"]
    padded = hdp != hd
    fused = Cc % 8 == 0 and attn.proj.bias is not None
    self._before_write(dz.data_ptr())
    if fused:
        has_g = br["gamma"] is not None
        self._op(ops, "fvit_branch_grad", gc_ptr, Cc, rows_c, Cc, P(br["gamma"]), br["s"], at["rs"], dz.data_ptr(), Cc,
                 br["w_alpha"], self.G(attn.proj.bias), at["u16"].data_ptr() if has_g else None, Cc, ("scal", 1),
                 self.G(blk.gamma1) if has_g else None)
    else:
        self._op(ops, "fvit_cast_scale_f16", gc_ptr, Cc, None, rows_c, Cc, P(br["gamma"]), br["s"], dz.data_ptr(), Cc,
                 at["rs"])
        if br["gamma"] is not None:
            self._op(ops, "fvit_colsum", gc_ptr, 0, Cc, None, at["u16"].data_ptr(), Cc, rows_c, Cc, None, ("scal", 1),
                     self.G(blk.gamma1), at["rs"])
    gWp = ("scr", self._scratch("dWproj_pad_c", Cc * Cp)) if padded else self.G(attn.proj.weight)
    self._linear_bwd(lin=attn.proj, w16=at["wp"].data_ptr(), ldw=at["ldp"], x16=at["ao"].data_ptr(), ldx=Cp,
                     dz16=dz.data_ptr(), lddz=Cc, rows=rows_c, n_out=Cc, k_in=Cp, br=br, gW=gWp, gW_ld=Cp,
                     dx16=dao.data_ptr(), lddx=Cp, flops_k=Cc, bias_done=fused)
    if padded:
        self._op(ops, "fvit_unpad_heads_f32", gWp, Cp, self.G(attn.proj.weight), Cc, Cc, Cp, hd, hdp, 0, 1, None)
        self._bwd_side.add(len(ops) - 1)   # fresh scratch in, parameter gradient out: leaf launch
    _emit_attn_core_bwd(self, at, dao, dqkv, B, n_ct)
    _bias_bwd(self, at["bias"])
    one = dict(gamma=None, s=None, inv_s=("scal", 2), w_alpha=("scal", 1))
    gWq, gbq = self.G(attn.qkv.weight), (self.G(attn.qkv.bias) if attn.qkv.bias is not None else None)
    if padded:
        gbq = ("scr", self._scratch("dbqkv_pad_c", 3 * Cp))
    self._linear_bwd(lin=attn.qkv, w16=at["wq"].data_ptr(), ldw=at["ldq"], x16=at["y16"].data_ptr(), ldx=Cc,
                     dz16=dqkv.data_ptr(), lddz=3 * Cp, rows=rows_c, n_out=3 * Cp, k_in=Cc, br=one, gW=gWq, gW_ld=Cc,
                     gW_row_map=_unpad_row_map(self, 3 * h, hd, hdp) if padded else None,
                     bias_to=gbq, dx16=dy.data_ptr(), lddx=Cc)
    if padded:
        if attn.qkv.bias is not None:
            self._op(ops, "fvit_unpad_heads_f32", gbq, 1, self.G(attn.qkv.bias), 1, 3 * Cp, 1, hd, hdp, 1, 0, None)
            self._bwd_side.add(len(ops) - 1)   # fresh scratch in, parameter gradient out: leaf launch
    # hat_norm1: rows r of the raster buffer = level rows ctr0 + r; gradient at (ct + hat_pe) first
    # accumulated in place (identity map), summed over images for hat_pos_embed, then moved to the
    # xs carrier rows it was gathered from
    ln = blk.hat_norm1
    self._op(ops, "fvit_ln_bwd", dy.data_ptr(), Cc, None, c["xh1"].data_ptr(), Cc, c["rs1"].data_ptr(),
             ln.weight.data_ptr(), rows_c, Cc, gc_ptr, Cc, None, 1, 0, ("scal", 1), self.G(ln.weight), self.G(ln.bias))
    if sv["hat_pe"] is not None:
        self._op(ops, "fvit_group_sum", gc_ptr, Cc, B, n_ct, 0, Cc, ("scal", 1), sv["hat_pe"]["dout"])
        _posemb_bwd(self, sv["hat_pe"])
    self._op(ops, "fvit_scatter_add_rows", gc_ptr, Cc, g_ptr, Cc, tl["ct_gather"].data_ptr(), rows_c, Cc)


class TokenLevelEmitters:
    """TrainPlan mixin: launch-list emitters of the transformer levels, the head and the build drivers (the functions
    above take the plan as `self`; `_attn_bwd_entry` is a plain helper)."""
    _drop_buf = _drop_buf
    _gen_drop_masks = _gen_drop_masks
    _build_forward = _build_forward
    _posemb_train = _posemb_train
    _bias_train = _bias_train
    _attn_fwd_train = _attn_fwd_train
    _gemm_train_branch = _gemm_train_branch
    _emit_token_level_train = _emit_token_level_train
    _mlp_fwd_train = _mlp_fwd_train
    _emit_head_train = _emit_head_train
    _build_backward = _build_backward
    _emit_head_bwd = _emit_head_bwd
    _emit_attn_core_bwd = _emit_attn_core_bwd
    _posemb_bwd = _posemb_bwd
    _bias_bwd = _bias_bwd
    _mlp_bwd = _mlp_bwd
    _attn_bwd = _attn_bwd
    _unpad_row_map = _unpad_row_map
    _emit_token_level_bwd = _emit_token_level_bwd
    _attn_bwd_carrier = _attn_bwd_carrier
