"""GPU tests of the inference-side API rows: deploy mode (fv.py:263-269, 336-342) and the dense-prediction backbone
interface (downstream/object_detection/dino/models/dino/fastervit.py:686-846) — per-level outputs against the CPU
oracle's captures of the same points, per-level BatchNorm against torch."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _setup(case):
    import fastervit_b200 as F
    from oracle import fastervit_oracle as O
    g = torch.load(GOLDEN / f"{case}.pt", weights_only=False)
    model = F.create_model(g["entry"], drop_path_rate=0.0, **g["kwargs"]).eval()
    O.synth_fill_(model.state_dict(), g["seeds"]["w"])
    x = O.synth_input(g["eval"]["batch"], g["cfg"]["resolution"], g["seeds"]["x"], torch.float32)
    return g, model.cuda(), x.cuda()


@pytest.mark.parametrize("case", ["tiny_a", "tiny_ar"])
def test_forward_levels_match_oracle_captures(case):
    from oracle import fastervit_oracle as O
    g, model, x = _setup(case)
    with torch.no_grad():
        feats = model.forward_levels(x, (0, 1, 2, 3))
        logits = model(x)
    sd = {k: (v.detach().cpu().double() if v.is_floating_point() else v.detach().cpu()) for k, v in model.state_dict().items()}
    cap = {}
    ref_logits = O.forward(sd, g["cfg"], x.cpu().double(), capture=cap)
    assert ((logits.cpu().double() - ref_logits).abs().max() / ref_logits.abs().max()).item() < 1e-3
    for i, f in enumerate(feats):
        want = cap[f"levels.{i}.out"]
        assert tuple(f.shape) == tuple(want.shape), (i, f.shape, want.shape)
        err = ((f.cpu().double() - want).abs().max() / want.abs().max()).item()
        assert err < 5e-3, (i, err)


def test_backbone_interface():
    import fastervit_b200 as F
    from fastervit_b200.backbone import FasterViTBackbone
    from oracle import fastervit_oracle as O
    kw = dict(resolution=[160, 224], dim=16, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 5, 7])
    torch.manual_seed(0)
    bb = FasterViTBackbone("faster_vit_0_any_res", out_indices=(1, 2, 3), frozen_stages=0, **kw)
    O.synth_fill_(bb.state_dict(), 11)
    bb = bb.cuda().eval()
    assert not any(p.requires_grad for p in bb.body.patch_embed.parameters())
    x = O.synth_input(2, [160, 224], 3, torch.float32).cuda()
    mask = torch.zeros(2, 160, 224, dtype=torch.bool, device="cuda")
    mask[1, :, 200:] = True
    outs = bb.forward_raw(x)
    raw = bb.body.forward_levels(x, (1, 2, 3))
    assert [tuple(o.shape) for o in outs] == [(2, 32, 20, 28), (2, 64, 10, 14), (2, 128, 5, 7)]
    for idx, o, r in zip((1, 2, 3), outs, raw):
        bn = getattr(bb, f"norm{idx}")
        ref = torch.nn.functional.batch_norm(r, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
        assert (o - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    d = bb(x, mask)
    assert set(d) == {0, 1, 2} and d[2][1].shape == (2, 5, 7) and d[0][1][1, :, -1].all() and not d[0][1][0].any()
    # a second input size builds a second plan (dynamic H x W per batch; like the reference any-res model the carrier
    # grid is fixed at construction, so the size must give the same number of level-2 windows: 10 x 14 -> 2 x 3)
    x2 = O.synth_input(1, [152, 216], 4, torch.float32).cuda()
    assert [tuple(o.shape) for o in bb.forward_raw(x2)] == [(1, 32, 19, 27), (1, 64, 10, 14), (1, 128, 5, 7)]


def test_switch_to_deploy_freezes_positional_tables():
    g, model, x = _setup("tiny_a")
    with torch.no_grad():
        ref = model(x)
        model.switch_to_deploy()
        assert all(m.deploy for m in model.modules() if hasattr(m, "deploy"))
        same = model(x)
        assert torch.equal(ref, same)
        # the positional MLPs are out of the loop now: changing them must not change the output ...
        for n, p in model.named_parameters():
            if "cpb_mlp" in n:
                p.mul_(1.5)
        frozen = model(x)
        assert torch.equal(ref, frozen)
        # ... while a model that is not in deploy mode follows them
        for m in model.modules():
            if hasattr(m, "deploy"):
                m.deploy = False
        moved = model(x)
    assert (moved - ref).abs().max().item() > 1e-4 * ref.abs().max().item()
