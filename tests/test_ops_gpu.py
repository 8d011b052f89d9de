"""GPU unit tests of the small positional-embedding kernels against plain torch (fp64)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,D", [(49, 784), (16, 64), (169, 16), (49, 1568), (60, 256), (7, 70), (529, 8), (1, 96)])
@pytest.mark.parametrize("save_hidden", [True, False])
def test_cpb_mlp_matches_torch(P, D, save_hidden):
    """cpb_mlp of PosEmbMLPSwinv1D / PosEmbMLPSwinv2D (fv.py:223-225, 322-324): Linear(2,512)+ReLU+Linear(512,D).
    D >= 64 takes the channel-parallel kernel, smaller D the point-parallel one."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(P * 1000 + D)
    coords = (torch.rand(P, 2, generator=g) * 2 - 1).cuda()
    w0 = torch.randn(512, 2, generator=g).cuda()
    b0 = torch.randn(512, generator=g).cuda()
    w1 = (torch.randn(D, 512, generator=g) * 0.05).cuda()
    out = torch.full((P, D), float("nan"), device="cuda")
    hid = torch.full((P, 512), float("nan"), device="cuda") if save_hidden else None
    L.call("fvit_cpb_mlp_fwd", coords.data_ptr(), P, w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), D, out.data_ptr(),
           L.ptr(hid))
    torch.cuda.synchronize()
    h_ref = torch.relu(coords.double() @ w0.double().t() + b0.double())
    ref = h_ref @ w1.double().t()
    assert (out.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-6
    if save_hidden:
        assert (hid.double() - h_ref).abs().max().item() <= 1e-5 * h_ref.abs().max().item()


def test_cpb_mlp_kernel_variants_agree_bitwise():
    """The two kernels use the same summation order, so a D >= 64 call equals the concatenation of D < 64 calls."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(3)
    P, D = 49, 128
    coords = (torch.rand(P, 2, generator=g) * 2 - 1).cuda()
    w0, b0 = torch.randn(512, 2, generator=g).cuda(), torch.randn(512, generator=g).cuda()
    w1 = (torch.randn(D, 512, generator=g) * 0.05).cuda()
    wide = torch.empty(P, D, device="cuda")
    L.call("fvit_cpb_mlp_fwd", coords.data_ptr(), P, w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), D, wide.data_ptr(), None)
    parts = []
    for d0 in range(0, D, 32):
        o = torch.empty(P, 32, device="cuda")
        L.call("fvit_cpb_mlp_fwd", coords.data_ptr(), P, w0.data_ptr(), b0.data_ptr(), w1[d0:d0 + 32].data_ptr(), 32,
               o.data_ptr(), None)
        parts.append(o)
    assert torch.equal(wide, torch.cat(parts, dim=1))


@pytest.mark.parametrize("P,D", [(49, 784), (16, 64), (169, 16), (100, 96), (70, 1568), (7, 70), (130, 8)])
def test_cpb_mlp_backward_matches_autograd(P, D):
    """fvit_cpb_mlp_bwd accumulates d w0, d b0, d w1 (scaled by a device scalar) from d out; D >= 64 takes the blocked
    kernels (8 channels per CTA / all points per thread), smaller D the per-channel / per-point ones."""
    from fastervit_b200 import lib as L
    g = torch.Generator().manual_seed(P * 7 + D)
    coords = (torch.rand(P, 2, generator=g) * 2 - 1).cuda()
    w0 = torch.randn(512, 2, generator=g).cuda().requires_grad_(True)
    b0 = torch.randn(512, generator=g).cuda().requires_grad_(True)
    w1 = (torch.randn(D, 512, generator=g) * 0.05).cuda().requires_grad_(True)
    dout = torch.randn(P, D, generator=g).cuda()
    hid = torch.empty(P, 512, device="cuda")
    out = torch.empty(P, D, device="cuda")
    L.call("fvit_cpb_mlp_fwd", coords.data_ptr(), P, w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), D, out.data_ptr(),
           hid.data_ptr())
    scale = torch.tensor([0.5], device="cuda")
    dw0 = torch.full((512, 2), 1.0, device="cuda")      # the kernel accumulates
    db0 = torch.full((512,), 1.0, device="cuda")
    dw1 = torch.full((D, 512), 1.0, device="cuda")
    L.call("fvit_cpb_mlp_bwd", coords.data_ptr(), P, w1.data_ptr(), hid.data_ptr(), dout.data_ptr(), D, scale.data_ptr(),
           dw0.data_ptr(), db0.data_ptr(), dw1.data_ptr())
    ref = torch.relu(coords.double() @ w0.double().t() + b0.double()) @ w1.double().t()
    gw0, gb0, gw1 = torch.autograd.grad(ref, [w0, b0, w1], grad_outputs=dout.double() * 0.5)
    for got, want, name in ((dw0, gw0, "dw0"), (db0, gb0, "db0"), (dw1, gw1, "dw1")):
        err = ((got.double() - 1.0) - want.double()).abs().max().item()
        assert err <= 2e-5 * want.abs().max().item() + 1e-5, (name, err)
