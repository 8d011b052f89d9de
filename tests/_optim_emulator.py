"""TEST-ONLY host emulator of the fvit_optim_* entry points (include/fvit.h), used by tests/test_optim_cpu.py to
drive the *host logic* of fastervit_b200.optim (chunk / pointer / offset tables, hyper-parameter sync, state
adoption, EMA aliasing, call signatures) in the CPU container, where no kernel can run. It reads and writes host
memory through the raw addresses the product code passes, exactly as the kernels do with device addresses, and
validates every call against the ctypes signature table. It is not a fallback: the product never imports it, and
`-m gpu` tests exercise the real kernels."""
import ctypes
import math

import numpy as np

from fastervit_b200 import lib as L

calls: list[str] = []


def _f32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr)) if n else np.zeros(0, np.float32)


def _i64_at(ptr, i):
    return ctypes.c_int64.from_address(ptr + 8 * i).value


def _chunks(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_int32 * (4 * n)).from_address(ptr)).reshape(n, 4).tolist()


def _check_sig(name, args):
    sig = L._OP_SIGS[name]
    assert len(args) + 1 == len(sig), f"{name}: {len(args)} args + stream vs {len(sig)} in the ctypes table"
    for a, t in zip(args, sig):
        if t is L._P:
            assert a is None or isinstance(a, int), (name, a)
        elif t in (L._I, L._L):
            assert isinstance(a, int), (name, a)
        else:
            assert isinstance(a, float), (name, a)


def gather(chunks, n, seg_src, seg_off, flat):
    for seg, st, cnt, _ in _chunks(chunks, n):
        src = _f32(_i64_at(seg_src, seg) + 4 * st, cnt)
        _f32(flat + 4 * (_i64_at(seg_off, seg) + st), cnt)[:] = src


def sqnorm(chunks, n, seg_off, g, partials):
    out = _f32(partials, 2 * n)
    for c, (seg, st, cnt, _) in enumerate(_chunks(chunks, n)):
        v = _f32(g + 4 * (_i64_at(seg_off, seg) + st), cnt)
        with np.errstate(all="ignore"):
            out[2 * c] = np.float32((v.astype(np.float64) ** 2).sum())
        out[2 * c + 1] = np.float32((~np.isfinite(v)).sum())


def prepare(partials, n, grad_scale, found_inf, max_norm, clip_eps, b1, b2, scal):
    sc = _f32(scal, 8)
    pr = _f32(partials, 2 * n) if n else np.zeros(0, np.float32)
    s = float(pr[0::2].astype(np.float64).sum())
    bad = float(pr[1::2].sum())
    inv = 1.0 / float(_f32(grad_scale, 1)[0]) if grad_scale else 1.0
    skip = bad > 0 or not math.isfinite(s) or not math.isfinite(inv) or (found_inf and _f32(found_inf, 1)[0] != 0)
    norm = math.sqrt(s) * inv if math.isfinite(s) else float("inf")
    clip = min(1.0, max_norm / (norm + clip_eps)) if (max_norm > 0 and norm + clip_eps > 0) else 1.0
    sc[0], sc[1], sc[2] = norm, 1.0 if skip else 0.0, inv * clip
    if not skip:
        sc[3] += 1.0
        sc[4] = 1.0 - b1 ** float(sc[3])
        sc[5] = 1.0 - b2 ** float(sc[3])


def _ema_blend(seg_ema, seg, st, cnt, p, decay):
    if seg_ema:
        e = _f32(_i64_at(seg_ema, seg) + 4 * st, cnt)
        e[:] = e * np.float32(decay) + np.float32(1.0 - decay) * p


def adamw(chunks, n, seg_p, seg_off, seg_hp, g, m, v, b1, b2, eps, scal, seg_ema, ema_decay):
    sc = _f32(scal, 8)
    skip = sc[1] != 0
    if skip and not seg_ema:
        return
    for seg, st, cnt, _ in _chunks(chunks, n):
        p = _f32(_i64_at(seg_p, seg) + 4 * st, cnt)
        if not skip:
            fo = _i64_at(seg_off, seg) + st
            gg = _f32(g + 4 * fo, cnt) * sc[2]
            mm, vv = _f32(m + 4 * fo, cnt), _f32(v + 4 * fo, cnt)
            lr, wd = _f32(seg_hp + 8 * seg, 2)
            p *= np.float32(1.0) - lr * wd
            mm += (gg - mm) * np.float32(1.0 - b1)
            vv[:] = vv * np.float32(b2) + np.float32(1.0 - b2) * gg * gg
            denom = np.sqrt(vv) / np.sqrt(sc[5]) + np.float32(eps)
            p -= (lr / sc[4]) * (mm / denom)
        _ema_blend(seg_ema, seg, st, cnt, p, ema_decay)


def lamb1(chunks, n, seg_p, seg_off, seg_hp, g, u, m, v, b1, b2, eps, scal, seg_norms):
    sc = _f32(scal, 8)
    if sc[1] != 0:
        return
    for seg, st, cnt, _ in _chunks(chunks, n):
        p = _f32(_i64_at(seg_p, seg) + 4 * st, cnt)
        fo = _i64_at(seg_off, seg) + st
        gg = _f32(g + 4 * fo, cnt) * sc[2]
        mm, vv, uu = _f32(m + 4 * fo, cnt), _f32(v + 4 * fo, cnt), _f32(u + 4 * fo, cnt)
        wd = _f32(seg_hp + 8 * seg, 2)[1]
        mm[:] = mm * np.float32(b1) + np.float32(1.0 - b1) * gg
        vv[:] = vv * np.float32(b2) + np.float32(1.0 - b2) * gg * gg
        denom = np.sqrt(vv) / np.sqrt(sc[5]) + np.float32(eps)
        uu[:] = (mm / sc[4]) / denom + wd * p
        nr = _f32(seg_norms + 8 * seg, 2)
        nr[0] += np.float32((p.astype(np.float64) ** 2).sum())
        nr[1] += np.float32((uu.astype(np.float64) ** 2).sum())


def lamb2(chunks, n, seg_p, seg_off, seg_hp, u, seg_norms, trust_clip, always_adapt, scal, seg_ema, ema_decay):
    sc = _f32(scal, 8)
    skip = sc[1] != 0
    if skip and not seg_ema:
        return
    for seg, st, cnt, _ in _chunks(chunks, n):
        p = _f32(_i64_at(seg_p, seg) + 4 * st, cnt)
        if not skip:
            lr, wd = _f32(seg_hp + 8 * seg, 2)
            trust = 1.0
            if wd != 0 or always_adapt:
                wn, un = np.sqrt(_f32(seg_norms + 8 * seg, 2))
                if wn > 0 and un > 0:
                    trust = float(wn / un)
                if trust_clip:
                    trust = min(trust, 1.0)
            p -= np.float32(lr * trust) * _f32(u + 4 * (_i64_at(seg_off, seg) + st), cnt)
        _ema_blend(seg_ema, seg, st, cnt, p, ema_decay)


def ema(chunks, n, seg_ema, seg_src, decay):
    for seg, st, cnt, _ in _chunks(chunks, n):
        src = _f32(_i64_at(seg_src, seg) + 4 * st, cnt)
        _ema_blend(seg_ema, seg, st, cnt, src, decay)


_IMPL = {"fvit_optim_gather_f32": gather, "fvit_optim_sqnorm": sqnorm, "fvit_optim_prepare": prepare,
         "fvit_optim_adamw": adamw, "fvit_optim_lamb_stage1": lamb1, "fvit_optim_lamb_stage2": lamb2,
         "fvit_optim_ema": ema}


def call(name, *args):
    _check_sig(name, args)
    calls.append(name)
    _IMPL[name](*args)
