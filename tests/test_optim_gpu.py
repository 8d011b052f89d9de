"""GPU parity of the optimizer row (train.py:879-899): the fused AdamW / LAMB / clip / EMA kernels of
libfvit_sm100.so against the fp64 CPU oracle (oracle/optim_oracle.py, pinned against torch.optim.AdamW and
clip_grad_norm_ in tests/test_optim_cpu.py) and against torch's own CUDA optimizers.

Tolerances (fp32 kernels vs fp64 oracle): moments rtol 2e-5; the parameter *update* (p_new - p_old) within 2e-5 of
its largest element per tensor after several steps; the gradient norm 1e-5 relative."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(dim=32, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8])
# odd sizes on purpose: scalar tails, sub-chunk, multi-chunk (CHUNK = 16384) and non-multiple-of-4 tensors
SHAPES = [(1,), (3,), (5, 7), (1000,), (16385,), (129, 543), (4, 4, 4, 4), (70001,)]


def _params(seed, shapes=SHAPES):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g, dtype=torch.float64) for s in shapes]


def _close_update(p_new, p_old, ref_new, ref_old, tol=2e-5, steps=1):
    du = p_new.double().cpu() - p_old
    dr = ref_new - ref_old
    scale = dr.abs().max().item() + 1e-30
    # parameters are stored in fp32: every step rounds p to half an ulp of |p|
    ulp = 1.2e-7 * max(ref_new.abs().max().item(), ref_old.abs().max().item())
    assert (du - dr).abs().max().item() <= tol * scale + (steps + 1) * ulp, ((du - dr).abs().max().item(), scale)


def _mk(shapes, seed, flat: bool):
    """CUDA fp32 parameters + a function that installs gradients either as views of one flat buffer (the layout
    the backward pass produces) or as separate tensors (gather path)."""
    from fastervit_b200 import optim as FO
    p64 = _params(seed, shapes)
    ps = [torch.nn.Parameter(p.float().cuda()) for p in p64]
    p64 = [p.detach().double().cpu() for p in ps]   # fp32-rounded starting point for the oracle
    offs, n = FO.sequential_offsets([p.numel() for p in ps])

    def set_grads(gs64):
        if flat:
            buf = torch.zeros(n, device="cuda")
            for p, o, g in zip(ps, offs, gs64):
                v = buf[o:o + p.numel()].view_as(p)
                v.copy_(g.float())
                p.grad = v
        else:
            for p, g in zip(ps, gs64):
                p.grad = g.float().cuda()
        return [p.grad.detach().double().cpu() for p in ps]
    return ps, p64, set_grads


@pytest.mark.parametrize("flat", [True, False])
def test_adamw_matches_oracle(flat):
    from fastervit_b200 import optim as FO
    from oracle import optim_oracle as OO
    ps, p64, set_grads = _mk(SHAPES, 0, flat)
    wd = [0.0 if i % 3 == 0 else 0.05 for i in range(len(ps))]
    groups = [{"params": [p for p, w in zip(ps, wd) if w == 0.0], "weight_decay": 0.0},
              {"params": [p for p, w in zip(ps, wd) if w != 0.0], "weight_decay": 0.05}]
    opt = FO.FusedAdamW(groups, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    ms = [torch.zeros_like(p) for p in p64]
    vs = [torch.zeros_like(p) for p in p64]
    start = [p.clone() for p in p64]
    for step in range(1, 6):
        lr = 1e-2 / step
        for grp in opt.param_groups:
            grp["lr"] = lr
        gs = set_grads(_params(100 + step))
        opt.step()
        for p, g, m, v, w in zip(p64, gs, ms, vs, wd):
            OO.adamw_step(p, g, m, v, step=step, lr=lr, weight_decay=w)
    torch.cuda.synchronize()
    assert (opt._lay["gstage"] is None) == flat          # flat gradients are read in place
    assert float(opt.state[ps[0]]["step"]) == 5.0
    for p, q, q0, m, v in zip(ps, p64, start, ms, vs):
        _close_update(p.detach(), q0, q, q0, steps=5)
        torch.testing.assert_close(opt.state[p]["exp_avg"].double().cpu(), m, rtol=2e-5, atol=5e-7)   # |g| ~ 1
        torch.testing.assert_close(opt.state[p]["exp_avg_sq"].double().cpu(), v, rtol=2e-5, atol=1e-9)


def test_adamw_matches_torch_cuda_adamw_with_external_clipping():
    """The reference loop: backward -> clip_grad_norm_(5.0) -> optimizer.step() (train.py:889-894)."""
    from fastervit_b200 import optim as FO
    ps, _, set_grads = _mk(SHAPES, 1, True)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    mine = FO.FusedAdamW(ps, lr=3e-3, weight_decay=0.1)
    theirs = torch.optim.AdamW(ref, lr=3e-3, weight_decay=0.1)
    for step in range(4):
        set_grads([g * 3 for g in _params(200 + step)])
        for r, p in zip(ref, ps):
            r.grad = p.grad.detach().clone()
        n1 = torch.nn.utils.clip_grad_norm_(ps, 5.0)
        n2 = torch.nn.utils.clip_grad_norm_(ref, 5.0)
        assert torch.allclose(n1, n2)
        mine.step()
        theirs.step()
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p.detach(), r.detach(), rtol=1e-5, atol=2e-6)


def test_fused_clipping_equals_clip_grad_norm():
    from fastervit_b200 import optim as FO
    from oracle import optim_oracle as OO
    ps, p64, set_grads = _mk(SHAPES, 2, True)
    start = [p.clone() for p in p64]
    opt = FO.FusedAdamW(ps, lr=1e-2, weight_decay=0.0, max_grad_norm=0.7)
    ms = [torch.zeros_like(p) for p in p64]
    vs = [torch.zeros_like(p) for p in p64]
    for step in range(1, 4):
        gs = set_grads(_params(300 + step))
        opt.step()
        norm, coef = OO.clip_coef(gs, 0.7)
        assert coef < 1.0
        assert abs(opt.grad_norm.item() - norm) <= 1e-5 * norm
        for p, g, m, v in zip(p64, gs, ms, vs):
            OO.adamw_step(p, g * coef, m, v, step=step, lr=1e-2, weight_decay=0.0)
    for p, q, q0 in zip(ps, p64, start):
        _close_update(p.detach(), q0, q, q0, steps=3)


@pytest.mark.parametrize("trust_clip,always_adapt", [(False, False), (True, True)])
def test_lamb_matches_oracle(trust_clip, always_adapt):
    from fastervit_b200 import optim as FO
    from oracle import optim_oracle as OO
    ps, p64, set_grads = _mk(SHAPES, 3, True)
    wd = [0.0 if p.ndim <= 1 else 0.12 for p in ps]     # timm's filter: no decay (hence no trust ratio) for 1-D
    groups = [{"params": [p for p, w in zip(ps, wd) if w == 0.0], "weight_decay": 0.0},
              {"params": [p for p, w in zip(ps, wd) if w != 0.0], "weight_decay": 0.12}]
    order = groups[0]["params"] + groups[1]["params"]
    idx = [next(i for i, p in enumerate(ps) if p is q) for q in order]
    opt = FO.FusedLamb(groups, lr=5e-3, max_grad_norm=1.0, trust_clip=trust_clip, always_adapt=always_adapt)
    ms = [torch.zeros_like(p) for p in p64]
    vs = [torch.zeros_like(p) for p in p64]
    start = [p.clone() for p in p64]
    for step in range(1, 5):
        gs = set_grads(_params(400 + step))
        opt.step()
        OO.lamb_step([p64[i] for i in idx], [gs[i] for i in idx], [ms[i] for i in idx], [vs[i] for i in idx],
                     step=step, lr=5e-3, weight_decay=[wd[i] for i in idx], max_grad_norm=1.0,
                     trust_clip=trust_clip, always_adapt=always_adapt)
    for p, q, q0 in zip(ps, p64, start):
        _close_update(p.detach(), q0, q, q0, tol=5e-5, steps=4)


def test_grad_scaler_integration_and_overflow_skip():
    """torch.amp.GradScaler drives the optimizer through grad_scale / found_inf device scalars: a non-finite
    gradient skips the step (parameters, moments and the step counter untouched), like the reference's
    NativeScaler path (train.py:879-886)."""
    from fastervit_b200 import optim as FO
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.GELU(), torch.nn.Linear(53, 10)).cuda()
    twin = copy.deepcopy(net)
    mine = FO.FusedAdamW(net.parameters(), lr=1e-2, weight_decay=0.01)
    theirs = torch.optim.AdamW(twin.parameters(), lr=1e-2, weight_decay=0.01)
    s1 = torch.amp.GradScaler("cuda", init_scale=1024.0)
    s2 = torch.amp.GradScaler("cuda", init_scale=1024.0)
    x = torch.randn(16, 37, device="cuda")
    y = torch.randint(0, 10, (16,), device="cuda")
    for it in range(4):
        for model, opt, sc in ((net, mine, s1), (twin, theirs, s2)):
            opt.zero_grad()
            loss = torch.nn.functional.cross_entropy(model(x), y)
            sc.scale(loss).backward()
            if it == 2:   # poison one gradient: both optimizers must skip this step
                next(model.parameters()).grad[0, 0] = float("inf")
            sc.step(opt)
            sc.update()
        if it == 2:
            assert mine.found_inf_flag.item() == 1.0
    assert float(mine.state[next(net.parameters())]["step"]) == 3.0
    assert s1.get_scale() == s2.get_scale() == 512.0
    for p, r in zip(net.parameters(), twin.parameters()):
        torch.testing.assert_close(p.detach(), r.detach(), rtol=1e-5, atol=1e-6)


def test_state_dict_round_trip_continues_identically():
    from fastervit_b200 import optim as FO
    ps, _, set_grads = _mk(SHAPES[:5], 5, True)
    opt = FO.FusedAdamW(ps, lr=1e-2, weight_decay=0.05)
    for step in range(2):
        set_grads(_params(500 + step, SHAPES[:5]))
        opt.step()
    sd = copy.deepcopy(opt.state_dict())
    snap = [p.detach().clone() for p in ps]
    set_grads(_params(502, SHAPES[:5]))
    opt.step()
    want = [p.detach().clone() for p in ps]
    # fresh parameters + fresh optimizer resumed from the checkpoint
    ps2 = [torch.nn.Parameter(s.clone()) for s in snap]
    opt2 = FO.FusedAdamW(ps2, lr=1e-2, weight_decay=0.05)
    opt2.load_state_dict(sd)
    g = _params(502, SHAPES[:5])
    for p, gg in zip(ps2, g):
        p.grad = gg.float().cuda()
    opt2.step()
    assert float(opt2.state[ps2[0]]["step"]) == 3.0
    for a, b in zip(ps2, want):
        torch.testing.assert_close(a.detach(), b, rtol=1e-6, atol=1e-7)


def _tiny_model(seed):
    import fastervit_b200 as F
    from oracle import fastervit_oracle as O
    m = F.create_model("faster_vit_0_224", drop_path_rate=0.0, **TINY)
    O.synth_fill_(m.state_dict(), seed)
    return m.cuda()


def test_ema_matches_model_ema_v2_loop_and_fused_variant():
    from fastervit_b200 import optim as FO
    from oracle import optim_oracle as OO
    model = _tiny_model(7)
    ema = FO.FlatEma(model, decay=0.9)
    ref = {k: v.detach().double().cpu().clone() if v.is_floating_point() else v.detach().cpu().clone()
           for k, v in ema.module.state_dict().items()}
    # alias map of the reference loop: entries that share storage in the module share it in the oracle copy too
    seen = {}
    for k, v in ema.module.state_dict().items():
        if v.data_ptr() in seen:
            ref[k] = ref[seen[v.data_ptr()]]
        seen.setdefault(v.data_ptr(), k)
    for it in range(3):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.1 * (it + 1))
            for b in model.buffers():
                if b.is_floating_point():
                    b.mul_(1.1)
                else:
                    b.add_(3)
        ema.update(model)
        msd = model.state_dict()
        for k in ref:   # timm ModelEmaV2.update
            mv = msd[k].detach().cpu()
            if ref[k].is_floating_point():
                OO.ema_update(ref[k], mv.double(), 0.9)
            else:
                ref[k].copy_(0.9 * ref[k] + 0.1 * mv)
    for k, v in ema.module.state_dict().items():
        if v.is_floating_point():
            torch.testing.assert_close(v.double().cpu(), ref[k], rtol=1e-5, atol=1e-6, msg=k)
        else:
            assert torch.equal(v.cpu(), ref[k]), k
    ema.set(model)
    for (k, v), w in zip(ema.module.state_dict().items(), model.state_dict().values()):
        assert torch.equal(v, w), k


def test_training_step_reads_backward_gradients_in_place_and_fused_ema_equals_separate_ema():
    """model(x) -> loss.backward() -> optimizer.step(): the flat gradient buffer of the backward pass is consumed
    without a copy; the update equals the oracle's AdamW on those gradients; EMA fused into the update kernel
    equals a separate FlatEma.update after the step (same run: the backward's atomics make two runs differ in the
    last bits, which Adam's sign-like first step would amplify)."""
    from fastervit_b200 import optim as FO
    from oracle import optim_oracle as OO
    from oracle import fastervit_oracle as O
    model = _tiny_model(11).train()
    opt = FO.FusedAdamW(model, lr=1e-3, weight_decay=0.05, max_grad_norm=5.0)
    ema = FO.FlatEma(model, decay=0.99)
    ema2 = FO.FlatEma(model, decay=0.99)      # not attached: separate 12-byte pass
    opt.attach_ema(ema, model)
    x = O.synth_input(2, 224, 5, torch.float32).cuda()
    y = torch.tensor([3, 7], device="cuda")
    for it in range(2):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        before = {n: p.detach().double().cpu() for n, p in model.named_parameters()}
        grads = {n: p.grad.detach().double().cpu() for n, p in model.named_parameters()}
        opt.step()
        if it == 0:
            assert opt._lay["gstage"] is None, "backward gradients should be consumed in place"
            _, coef = OO.clip_coef(list(grads.values()), 5.0)
            for n, p in model.named_parameters():
                q = before[n].clone()
                wd = 0.0 if (p.ndim <= 1 or n.endswith(".bias")) else 0.05
                OO.adamw_step(q, grads[n] * coef, torch.zeros_like(q), torch.zeros_like(q), step=1, lr=1e-3,
                              weight_decay=wd)
                _close_update(p.detach(), before[n], q, before[n])
        ema.update(model)     # buffers + the second blend of the aliased tokenizer conv
        ema2.update(model)
    torch.cuda.synchronize()
    assert float(opt.state[next(model.parameters())]["step"]) == 2.0
    moved = 0
    for (k, a), b, w in zip(ema.module.state_dict().items(), ema2.module.state_dict().values(),
                            model.state_dict().values()):
        if a.is_floating_point():
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=k)
            moved += int(not torch.equal(a, w))
        else:
            assert torch.equal(a, b), k
    assert moved > 100   # the EMA lags the model: it was really blended, not copied
