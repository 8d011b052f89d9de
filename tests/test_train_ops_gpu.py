"""GPU unit tests of the HBM-bound training kernels (LayerNorm / BatchNorm backward, the normalise pass of a raw
convolution output, column reductions, carrier propagation and TokenInitializer backward) against fp64 torch —
autograd of the reference's own ops where there is one (fv.py has no backward code, train.py:879-896 relies on
autograd). Operands are the fp16 tensors the launch list really passes, so the tolerance is rounding of the
fp32 arithmetic only: 1e-4 relative for fp32 outputs, 2e-3 for fp16 outputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rel, what=""):
    ref = ref.double()
    d = (got.double().cpu() - ref.cpu()).abs().max().item()
    assert d <= rel * max(ref.abs().max().item(), 1e-6), (what, d, ref.abs().max().item())


_KEEP = []   # device copies handed to the C ABI by raw pointer must outlive the launch (and the allocator's reuse)


def _scalar(v):
    t = torch.tensor([v], dtype=torch.float32, device="cuda")
    _KEEP.append(t)
    return t


def _dev(t):
    """device pointer of a CUDA copy of t that stays alive until the end of the test module"""
    c = t.cuda().contiguous()
    _KEEP.append(c)
    return c.data_ptr()


@pytest.mark.parametrize("rows,C", [(212, 256), (53, 784), (1000, 64), (96, 1568), (37, 24)])
@pytest.mark.parametrize("mapped", [False, True])
def test_ln_bwd_matches_autograd(rows, C, mapped):
    """fvit_ln_bwd vs autograd of nn.LayerNorm on the same (fp16-rounded) normalised input: gradient into the fp32
    residual-stream buffer (accumulating, optionally scattered through in_map), dgamma, dbeta."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(rows * 7 + C)
    x = torch.randn(rows, C, generator=g, dtype=torch.float64) * 1.7 + 0.3
    gamma = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5)
    mu, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    xhat16 = ((x - mu) * rstd).half()
    dy16 = (torch.randn(rows, C, generator=g) * 0.5).half()
    # mapped: every 5th row's gradient is routed to a row of a second region (the carrier slots of a window go
    # back to the raster-ordered carrier rows, engine.py norm1_gather) and the slot itself is cleared
    moved = (torch.arange(rows) % 5 == 0) if mapped else torch.zeros(rows, dtype=torch.bool)
    dest = torch.arange(rows)
    dest[moved] = rows + torch.arange(int(moved.sum()))
    total = rows + int(moved.sum())
    g_in = torch.randn(total, C, generator=g, dtype=torch.float32)
    # reference: y = xhat * gamma + beta, xhat = normalise(x); d/dx through the stored xhat and rstd
    xh, dy = xhat16.double(), dy16.double()
    gd = gamma * dy
    dx = rstd * (gd - gd.mean(1, keepdim=True) - xh * (gd * xh).mean(1, keepdim=True))
    want = g_in.double().clone()
    want[dest] = g_in[:rows].double() + dx    # gv = g[r] + dx[r] is written to row in_map[r]
    want[:rows][moved] = 0.0                   # clear_moved
    scal = 0.25
    gbuf = g_in.clone().cuda()
    dgam = torch.zeros(C, device="cuda")
    dbet = torch.zeros(C, device="cuda")
    in_map = dest.to(torch.int32).cuda() if mapped else None
    L.call("fvit_ln_bwd", _dev(dy16), C, None, _dev(xhat16), C, _dev(rstd.float().flatten()),
           _dev(gamma.float()), rows, C, gbuf.data_ptr(), C, L.ptr(in_map), 1, 1 if mapped else 0,
           _scalar(scal).data_ptr(), dgam.data_ptr(), dbet.data_ptr())
    torch.cuda.synchronize()
    _close(gbuf, want, 2e-5, "g")
    _close(dgam, scal * (dy * xh).sum(0), 1e-4, "dgamma")
    _close(dbet, scal * dy.sum(0), 1e-4, "dbeta")
    # and the formula itself against autograd (unrounded xhat): fp16 rounding of xhat / dy only
    xr = x.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (C,), gamma, torch.zeros(C, dtype=torch.float64), 1e-5)
    y.backward(dy)
    _close(dx, xr.grad, 3e-3, "formula")


@pytest.mark.parametrize("rows,C,act", [(784, 64, 0), (300, 200, 1), (1568, 392, 0), (50, 16, 1), (900, 196, 0), (333, 36, 1)])
@pytest.mark.parametrize("g_is_f16", [0, 1])
def test_bn_bwd_matches_autograd(rows, C, act, g_is_f16):
    """fvit_bn_bwd vs autograd of train-mode BatchNorm (+ReLU) over a list of rows of a raw convolution output; fp16
    rows are padded to a multiple of 8 columns like the level buffers (fv4: C = 196 -> 200)."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(rows + C + act)
    pad = 5   # the row lists address a larger, bordered buffer
    ld = (C + 7) // 8 * 8
    raw16 = torch.zeros(rows + pad, ld, dtype=torch.float16)
    raw16[:, :C] = (torch.randn(rows + pad, C, generator=g) * 2 + 0.5).half()
    r_rows = (torch.randperm(rows + pad, generator=g)[:rows]).to(torch.int32)
    gin = torch.randn(rows + pad, C, generator=g) * 0.3
    if g_is_f16:
        g16 = torch.zeros(rows + pad, ld, dtype=torch.float16)
        g16[:, :C] = gin.half()
        gin_dev, ldg, gin = g16, ld, g16[:, :C]
    else:
        gin_dev, ldg = gin, C
    g_rows = (torch.randperm(rows + pad, generator=g)[:rows]).to(torch.int32)
    o_rows = (torch.randperm(rows + pad, generator=g)[:rows]).to(torch.int32)
    w = torch.rand(C, generator=g) + 0.5
    b = torch.randn(C, generator=g) * 0.2
    colmul = torch.rand(C, generator=g) + 0.5
    rsc = (torch.rand(rows + pad, generator=g) > 0.3).float() * 1.25
    xr = raw16[r_rows.long(), :C].double().requires_grad_(True)
    mean, var = xr.mean(0), xr.var(0, unbiased=False)
    eps = 1e-5
    rstd = (var + eps).rsqrt()
    wd = w.double().clone().requires_grad_(True)   # separate leaves for dw / db
    bd = b.double().clone().requires_grad_(True)
    y2 = (xr - mean) * rstd * wd + bd
    if act == 1:
        y2 = torch.relu(y2)
    upstream = gin[g_rows.long()].double() * colmul.double() * rsc[g_rows.long()].double().view(-1, 1)
    y2.backward(upstream)
    scal = 0.5
    out16 = torch.zeros(rows + pad, ld, dtype=torch.float16, device="cuda")
    s1, s2 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dw, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    L.call("fvit_bn_bwd", _dev(gin_dev), g_is_f16, ldg, _dev(g_rows), _dev(raw16), ld,
           _dev(r_rows), rows, C, _dev(mean.detach().float()), _dev(rstd.detach().float()),
           _dev(w), _dev(b), act, _dev(colmul), s1.data_ptr(), s2.data_ptr(),
           _scalar(scal).data_ptr(), out16.data_ptr(), ld, _dev(o_rows), dw.data_ptr(), db.data_ptr(),
           _dev(rsc))
    torch.cuda.synchronize()
    _close(out16[o_rows.long().cuda()][:, :C], xr.grad, 2e-3, "dx")
    _close(dw, scal * wd.grad, 2e-4, "dw")
    _close(db, scal * bd.grad, 2e-4, "db")


@pytest.mark.parametrize("rows,C,act", [(640, 64, 2), (333, 200, 1), (100, 392, 0), (500, 196, 2), (77, 36, 1)])
def test_affine_rows_matches_torch(rows, C, act):
    """fvit_affine_rows: y = act(x16*scale + shift) * row_scale (+ resid) over a row list, fp32 and fp16 outputs.
    fp16 rows are padded to a multiple of 8 columns like the level buffers (fv4: C = 196 -> 200)."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(rows + C)
    total = rows + 7
    ld = (C + 7) // 8 * 8
    x16 = torch.zeros(total, ld, dtype=torch.float16)
    x16[:, :C] = (torch.randn(total, C, generator=g) * 1.5).half()
    rowsel = torch.randperm(total, generator=g)[:rows].to(torch.int32)
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    resid = torch.randn(total, C, generator=g)
    rsc = (torch.rand(total, generator=g) > 0.3).float() / 0.7
    v = x16[:, :C].double() * scale.double() + shift.double()
    if act == 1:
        v = torch.relu(v)
    elif act == 2:
        v = torch.nn.functional.gelu(v)
    ref = v * rsc.double().view(-1, 1) + resid.double()
    o32 = torch.zeros(total, C, device="cuda")
    o16 = torch.full((total, ld), 3.0, dtype=torch.float16, device="cuda")
    L.call("fvit_affine_rows", _dev(x16), ld, _dev(rowsel), rows, C, _dev(scale), _dev(shift), act, _dev(resid), C,
           o32.data_ptr(), C, o16.data_ptr(), ld, _dev(rsc))
    torch.cuda.synchronize()
    sel = rowsel.long()
    _close(o32[sel.cuda()], ref[sel], 2e-6 if act != 2 else 2e-6 + 3e-7, "out32")   # A-S erf: |error| < 1.5e-7 absolute
    _close(o16[sel.cuda()][:, :C], ref[sel], 1e-3, "out16")
    if ld > C:
        assert o16[sel.cuda()][:, C:].abs().max().item() == 0.0      # padding columns are written as zeros
    untouched = torch.ones(total, dtype=torch.bool)
    untouched[sel] = False
    assert o32[untouched.cuda()].abs().max().item() == 0.0


@pytest.mark.parametrize("a_is_f16", [0, 1])
def test_colsum_matches_torch(a_is_f16):
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(11 + a_is_f16)
    rows, C, total = 500, 136, 520
    a = torch.randn(total, C, generator=g)
    a = a.half() if a_is_f16 else a
    b16 = torch.randn(rows, C, generator=g).half()
    a_rows = torch.randperm(total, generator=g)[:rows].to(torch.int32)
    colmul = torch.rand(C, generator=g) + 0.5
    rsc = torch.rand(rows, generator=g)
    out = torch.full((C,), 2.0, device="cuda")
    L.call("fvit_colsum", _dev(a), a_is_f16, C, _dev(a_rows), _dev(b16), C, rows, C,
           _dev(colmul), _scalar(0.5).data_ptr(), out.data_ptr(), _dev(rsc))
    torch.cuda.synchronize()
    ref = 2.0 + 0.5 * colmul.double() * (a[a_rows.long()].double() * b16.double() * rsc.double().view(-1, 1)).sum(0)
    _close(out, ref, 1e-4, "colsum")


def test_propagate_bwd_matches_autograd():
    """fv.py:697-700 backward: g[src[r]] += gamma * g[r]; dgamma += scalar * sum_r g[r] * xs[src[r]]."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(5)
    nW, S, ncw, C = 6, 53, 4, 64
    rows = nW * S
    r = torch.arange(rows)
    t = r % S - ncw
    src = torch.where(t >= 0, (r // S) * S + (t % ncw), torch.full_like(r, -1)).to(torch.int32)
    xs = torch.randn(rows, C, generator=g)
    gbuf = torch.randn(rows, C, generator=g)
    gamma = torch.rand(C, generator=g) * 1e-2
    xd = xs.double().requires_grad_(True)
    gm = gamma.double().requires_grad_(True)
    valid = src >= 0
    y = xd.clone()
    y[valid] = xd[valid] + gm * xd[src[valid].long()]
    y.backward(gbuf.double())
    gb = gbuf.clone().cuda()
    dgam = torch.zeros(C, device="cuda")
    L.call("fvit_propagate_bwd", gb.data_ptr(), C, _dev(xs), C, _dev(src), rows, C,
           _dev(gamma), _scalar(2.0).data_ptr(), dgam.data_ptr())
    torch.cuda.synchronize()
    _close(gb, xd.grad, 1e-5, "g")
    _close(dgam, 2.0 * gm.grad, 1e-4, "dgamma")


def test_token_init_bwd_matches_autograd():
    """TokenInitializer backward (fv.py:733-738): depthwise 3x3 conv (+bias) then AvgPool(k5, s3) of a 14 x 14 map
    into a 4 x 4 carrier grid; carrier-row gradients -> pixel-row gradients, dw, dbias."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(9)
    B, Hp, Wp, C, oh, ow, kh, kw, sh, sw = 2, 14, 14, 48, 4, 4, 5, 5, 3, 3
    npix = B * Hp * Wp
    x = torch.randn(B, C, Hp, Wp, generator=g)
    x16 = x.permute(0, 2, 3, 1).reshape(npix, C).half()
    w = torch.randn(C, 1, 3, 3, generator=g) * 0.3
    gct = torch.randn(B, oh * ow, C, generator=g)
    xd = x16.double().view(B, Hp, Wp, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    wd = w.double().clone().requires_grad_(True)
    bd = torch.zeros(C, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(xd, wd, bd, padding=1, groups=C)
    y = torch.nn.functional.avg_pool2d(y, (kh, kw), (sh, sw))
    assert y.shape[-2:] == (oh, ow)
    y.backward(gct.double().view(B, oh, ow, C).permute(0, 3, 1, 2))
    gbuf = torch.zeros(npix + B * oh * ow, C)
    gpix0 = torch.randn(npix, C, generator=g)
    gbuf[:npix] = gpix0
    gbuf[npix:] = gct.reshape(-1, C)
    gbuf = gbuf.cuda()
    pix_map = torch.arange(npix, dtype=torch.int32).cuda()
    ct_rows = (npix + torch.arange(B * oh * ow)).to(torch.int32).cuda()
    dw = torch.zeros(C, 9, device="cuda")
    db = torch.zeros(C, device="cuda")
    L.call("fvit_token_init_bwd", gbuf.data_ptr(), C, _dev(x16), C, pix_map.data_ptr(), ct_rows.data_ptr(), B, Hp, Wp,
           C, _dev(w), kh, kw, sh, sw, oh, ow, _scalar(0.5).data_ptr(), gbuf.data_ptr(), C, dw.data_ptr(),
           db.data_ptr())
    torch.cuda.synchronize()
    want_pix = gpix0.double() + xd.grad.permute(0, 2, 3, 1).reshape(npix, C)
    _close(gbuf[:npix], want_pix, 1e-5, "gx")
    _close(dw, 0.5 * wd.grad.view(C, 9), 1e-4, "dw")
    _close(db, 0.5 * bd.grad, 1e-4, "dbias")
