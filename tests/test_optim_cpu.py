"""CPU tests of the optimizer row: the oracle pinned against torch itself, and the host-side layout
logic of fastervit_b200.optim (no kernels are launched here)."""
import copy
import sys
from pathlib import Path

import pytest
import torch

import fastervit_b200 as F
from fastervit_b200 import optim as FO
from fastervit_b200.lib import FvitError
from oracle import optim_oracle as OO

sys.path.insert(0, str(Path(__file__).resolve().parent))   # tests/_optim_emulator.py
TINY = dict(dim=32, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8])


def _rand_set(seed, shapes, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g, dtype=dtype) for s in shapes]


SHAPES = [(7,), (3, 5), (64, 9), (1,), (130, 33)]


def test_oracle_adamw_is_torch_adamw():
    """Pin: oracle.adamw_step == torch.optim.AdamW (fp64, 6 steps, two weight-decay groups, changing lr)."""
    ps = _rand_set(0, SHAPES)
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.AdamW([{"params": ref[:2], "weight_decay": 0.0}, {"params": ref[2:], "weight_decay": 0.05}],
                            lr=1e-2, betas=(0.9, 0.999), eps=1e-8, foreach=False)
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    for step in range(1, 7):
        lr = 1e-2 * (1 + 0.1 * step)
        for grp in opt.param_groups:
            grp["lr"] = lr
        gs = _rand_set(100 + step, SHAPES)
        for r, g in zip(ref, gs):
            r.grad = g.clone()
        opt.step()
        for i, (p, g, m, v) in enumerate(zip(ps, gs, ms, vs)):
            OO.adamw_step(p, g, m, v, step=step, lr=lr, weight_decay=0.0 if i < 2 else 0.05)
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p, r.detach(), rtol=1e-13, atol=1e-15)


def test_oracle_clip_is_torch_clip_grad_norm():
    gs = _rand_set(3, SHAPES)
    ref = [torch.nn.Parameter(torch.zeros_like(g)) for g in gs]
    for r, g in zip(ref, gs):
        r.grad = g.clone()
    total = torch.nn.utils.clip_grad_norm_(ref, max_norm=0.5)
    norm, coef = OO.clip_coef(gs, 0.5)
    assert abs(norm - float(total)) < 1e-12 * norm
    for r, g in zip(ref, gs):
        torch.testing.assert_close(r.grad, g * coef, rtol=1e-13, atol=0)
    assert OO.clip_coef(gs, 1e9)[1] == 1.0 and OO.clip_coef(gs, 0.0)[1] == 1.0


def test_oracle_lamb_trust_ratio_properties():
    """The LAMB restatement (unpinned against timm, see oracle/optim_oracle.py) obeys the published update rule:
    with weight decay the step length is lr * ||p|| per tensor, without it LAMB degenerates to Adam's direction."""
    ps = _rand_set(1, SHAPES)
    p0 = [p.clone() for p in ps]
    gs = _rand_set(2, SHAPES)
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    wd = [0.1] * len(ps)
    OO.lamb_step(ps, gs, ms, vs, step=1, lr=0.01, weight_decay=wd, max_grad_norm=None)
    for p, q in zip(ps, p0):
        assert abs(float((p - q).norm()) / (0.01 * float(q.norm())) - 1.0) < 1e-9
    # no decay: update = sign-like Adam direction m_hat / (sqrt(v_hat) + eps), |.| ~ 1 at step 1
    ps2 = [q.clone() for q in p0]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    OO.lamb_step(ps2, gs, ms, vs, step=1, lr=0.01, weight_decay=[0.0] * len(ps), max_grad_norm=None)
    for p, q, g in zip(ps2, p0, gs):
        torch.testing.assert_close(p, q - 0.01 * g / (g.abs() + 1e-6), rtol=1e-9, atol=1e-12)
    # global clipping: scaling all gradients up by 10 changes nothing once the norm exceeds max_grad_norm
    a = [q.clone() for q in p0]
    b = [q.clone() for q in p0]
    for tgt, scale in ((a, 1.0), (b, 10.0)):
        ms = [torch.zeros_like(p) for p in ps]
        vs = [torch.zeros_like(p) for p in ps]
        OO.lamb_step(tgt, [g * scale for g in gs], ms, vs, step=1, lr=0.01, weight_decay=wd, max_grad_norm=1.0)
    for x, y in zip(a, b):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-9)


def test_oracle_ema_is_model_ema_v2_loop():
    m = F.create_model("faster_vit_0_224", **TINY)
    e = copy.deepcopy(m)
    for p in m.parameters():
        p.data.add_(0.5)
    tok = e.levels[2].global_tokenizer.pos_embed.weight.clone()
    tgt = m.levels[2].global_tokenizer.pos_embed.weight
    with torch.no_grad():
        for ev, mv in zip(e.state_dict().values(), m.state_dict().values()):   # timm ModelEmaV2.update
            if ev.is_floating_point():
                OO.ema_update(ev, mv, 0.9)
    # the aliased depthwise conv appears under two state_dict keys and is therefore blended twice
    once = 0.9 * tok + 0.1 * tgt
    torch.testing.assert_close(e.levels[2].global_tokenizer.pos_embed.weight.detach(), 0.9 * once + 0.1 * tgt)


@pytest.mark.parametrize("chunk", [4, 64, 16384])
def test_chunk_table_covers_every_active_element_once(chunk):
    numels = [5, 0, 64, 65, 40000, 1, 16384]
    active = [True, True, False, True, True, True, True]
    tab = FO.build_chunks(numels, chunk, active)
    cover = [torch.zeros(n, dtype=torch.int32) for n in numels]
    for seg, start, count, pad in tab.tolist():
        assert pad == 0 and start % 4 == 0 and 0 < count <= chunk
        cover[seg][start:start + count] += 1
    for n, c, a in zip(numels, cover, active):
        assert bool((c == (1 if a else 0)).all())


def test_sequential_offsets_equal_the_backward_pass_layout():
    """optim.sequential_offsets must reproduce TrainPlan's flat gradient layout (engine_train._setup_train): that
    is what lets the optimizer read the gradients in place."""
    m = F.create_model("faster_vit_0_224", **TINY)
    ps = list(m.parameters())
    offs, n = FO.sequential_offsets([p.numel() for p in ps])
    ref, k = [], 0
    for p in ps:
        ref.append(k)
        k += (p.numel() + 63) // 64 * 64
    assert offs == ref and n == k
    flat = torch.zeros(n)
    views = [flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, ps)]
    assert FO.shared_storage_offsets(views) == offs
    assert FO.shared_storage_offsets([torch.zeros(3), torch.zeros(4)]) is None          # two storages
    assert FO.shared_storage_offsets([flat[0:10], flat[5:15]]) is None                    # overlap
    assert FO.shared_storage_offsets([flat[0:10], None]) is None
    assert FO.shared_storage_offsets([flat[0:10].double()]) is None


def test_param_groups_follow_timm_weight_decay_filter():
    m = F.create_model("faster_vit_0_224", **TINY)
    groups = FO.param_groups_weight_decay(m, 0.05)
    no_decay, decay = groups
    assert no_decay["weight_decay"] == 0.0 and decay["weight_decay"] == 0.05
    assert len(no_decay["params"]) + len(decay["params"]) == len(list(m.parameters()))
    assert all(p.ndim > 1 for p in decay["params"])
    names = {id(p): n for n, p in m.named_parameters()}
    assert all(p.ndim <= 1 or names[id(p)].endswith(".bias") for p in no_decay["params"])


def test_optimizer_is_a_torch_optimizer_and_has_no_cpu_path():
    m = F.create_model("faster_vit_0_224", **TINY)
    for cls in (FO.FusedAdamW, FO.FusedLamb):
        opt = cls(m, lr=1e-3, weight_decay=0.05)
        assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 2
        assert opt._step_supports_amp_scaling
        sd = opt.state_dict()           # before the first step: plain torch behaviour
        assert sd["state"] == {} and len(sd["param_groups"]) == 2
        for p in m.parameters():
            p.grad = torch.zeros_like(p)
        with pytest.raises(FvitError):
            opt.step()
        opt.zero_grad()
    ema = FO.FlatEma(m, decay=0.99)
    assert not ema.module.training and ema.module is not m
    with pytest.raises(FvitError):
        ema.update(m)


# ------------------------------------------------------------------------------------ host logic under the emulator
@pytest.fixture
def emu(monkeypatch):
    """Route the fvit_optim_* calls to tests/_optim_emulator.py (host memory) and lift the CUDA-only guards, so the
    product's table / offset / state logic runs end to end in the CPU container."""
    import contextlib
    import _optim_emulator as E
    from fastervit_b200 import lib as L
    E.calls.clear()
    monkeypatch.setattr(L, "call", E.call)
    monkeypatch.setattr(FO, "_require_cuda", lambda t, what: None)
    monkeypatch.setattr(FO, "_device_guard", lambda dev: contextlib.nullcontext())
    return E


def _mk_params(seed, flat):
    ps = [torch.nn.Parameter(p.float()) for p in _rand_set(seed, SHAPES + [(16385,), (4, 4099)])]
    offs, n = FO.sequential_offsets([p.numel() for p in ps])

    def set_grads(seed):
        gs = _rand_set(seed, [tuple(p.shape) for p in ps])
        if flat:
            buf = torch.zeros(n)
            for p, o, g in zip(ps, offs, gs):
                v = buf[o:o + p.numel()].view_as(p)
                v.copy_(g.float())
                p.grad = v
        else:
            for p, g in zip(ps, gs):
                p.grad = g.float()
        return [p.grad.double() for p in ps]
    return ps, set_grads


@pytest.mark.parametrize("flat", [True, False])
def test_host_logic_adamw_against_oracle(emu, flat):
    ps, set_grads = _mk_params(0, flat)
    p64 = [p.detach().double().clone() for p in ps]
    groups = [{"params": ps[1::2], "weight_decay": 0.0}, {"params": ps[0::2], "weight_decay": 0.05}]
    order = ps[1::2] + ps[0::2]
    wd = {id(p): w for g, w in ((groups[0], 0.0), (groups[1], 0.05)) for p in g["params"]}
    opt = FO.FusedAdamW(groups, lr=1e-2, max_grad_norm=2.0)
    ms = [torch.zeros_like(p) for p in p64]
    vs = [torch.zeros_like(p) for p in p64]
    for step in range(1, 4):
        for grp in opt.param_groups:
            grp["lr"] = 1e-2 / step
        gs = set_grads(10 + step)
        opt.step()
        _, coef = OO.clip_coef(gs, 2.0)
        for p, q, g, m, v in zip(ps, p64, gs, ms, vs):
            OO.adamw_step(q, g * coef, m, v, step=step, lr=1e-2 / step, weight_decay=wd[id(p)])
    assert (opt._lay["gstage"] is None) == flat
    assert ("fvit_optim_gather_f32" in emu.calls) == (not flat)
    assert emu.calls.count("fvit_optim_adamw") == 3 and emu.calls.count("fvit_optim_sqnorm") == 3
    assert float(opt.state[order[0]]["step"]) == 3.0
    for p, q, m in zip(ps, p64, ms):
        torch.testing.assert_close(p.detach().double(), q, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(opt.state[p]["exp_avg"].double(), m, rtol=1e-5, atol=1e-7)


def test_host_logic_lamb_partial_grads_and_state_dict(emu):
    ps, set_grads = _mk_params(1, True)
    p64 = [p.detach().double().clone() for p in ps]
    wd = [0.0 if p.ndim <= 1 else 0.1 for p in ps]
    opt = FO.FusedLamb([{"params": [p], "weight_decay": w} for p, w in zip(ps, wd)], lr=5e-3, max_grad_norm=1.0)
    ms = [torch.zeros_like(p) for p in p64]
    vs = [torch.zeros_like(p) for p in p64]
    for step in range(1, 4):
        gs = set_grads(20 + step)
        OO.lamb_step(p64, gs, ms, vs, step=step, lr=5e-3, weight_decay=wd, max_grad_norm=1.0)
        opt.step()
    for p, q in zip(ps, p64):
        torch.testing.assert_close(p.detach().double(), q, rtol=2e-5, atol=2e-6)
    # a parameter without a gradient is skipped for that step (torch semantics): its value and moments stay
    sd = copy.deepcopy(opt.state_dict())
    snap = [p.detach().clone() for p in ps]
    set_grads(30)
    ps[2].grad = None
    keep_m = opt.state[ps[2]]["exp_avg"].clone()
    opt.step()
    assert torch.equal(ps[2].detach(), snap[2]) and torch.equal(opt.state[ps[2]]["exp_avg"], keep_m)
    assert not torch.equal(ps[3].detach(), snap[3])
    # resume from the checkpoint into a fresh optimizer with a different gradient layout
    ps2 = [torch.nn.Parameter(s.clone()) for s in snap]
    opt2 = FO.FusedLamb([{"params": [p], "weight_decay": w} for p, w in zip(ps2, wd)], lr=5e-3, max_grad_norm=1.0)
    opt2.load_state_dict(sd)
    for p2, p in zip(ps2, ps):
        p2.grad = None if p.grad is None else p.grad.clone()
    opt2.step()
    assert float(opt2.state[ps2[0]]["step"]) == 4.0
    for a, b in zip(ps2, ps):
        torch.testing.assert_close(a.detach(), b.detach(), rtol=1e-6, atol=1e-7)


def test_host_logic_overflow_skips_step_but_not_fused_ema(emu):
    m = F.create_model("faster_vit_0_224", **TINY)
    opt = FO.FusedAdamW(m, lr=1e-3, weight_decay=0.05, max_grad_norm=5.0)
    ema = FO.FlatEma(m, decay=0.9)
    plain = FO.FlatEma(m, decay=0.9)
    opt.attach_ema(ema, m)
    g = torch.Generator().manual_seed(5)
    for it in range(3):
        for p in m.parameters():
            p.grad = torch.randn(p.shape, generator=g) * 0.01
        if it == 1:
            next(m.parameters()).grad.view(-1)[3] = float("nan")
        before = [p.detach().clone() for p in m.parameters()]
        with torch.no_grad():
            for b in m.buffers():
                if b.dtype == torch.float32:
                    b.add_(0.25)
        opt.step()
        ema.update(m)
        plain.update(m)
        changed = any(not torch.equal(a, p.detach()) for a, p in zip(before, m.parameters()))
        assert changed == (it != 1)
        assert opt.found_inf_flag.item() == (1.0 if it == 1 else 0.0)
    assert float(opt.state[next(m.parameters())]["step"]) == 2.0
    # fused (parameters blended inside the update kernel, also on the skipped step) == separate ModelEmaV2 pass,
    # including the tokenizer's depthwise conv that state_dict() lists twice
    for (k, a), b in zip(ema.module.state_dict().items(), plain.module.state_dict().values()):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=k)
    tok = ema.module.levels[2].global_tokenizer.pos_embed.weight
    assert not torch.equal(tok, m.levels[2].global_tokenizer.pos_embed.weight)
