"""Model configurations used by the oracle-side tests and golden generator (test infrastructure).

`REF_KWARGS[name]` are the kwargs passed to the reference's create_model; `cfg_of(name)` is the oracle
cfg dict. The full-size entries restate the entrypoint defaults (fv.py:977-1166, fvar.py:1007-1019);
the `tiny_*` entries are reduced-width models that exercise every code path in seconds on a CPU.
"""
from __future__ import annotations

CASES = {
    # name: (reference entrypoint, override kwargs, oracle cfg)
    "fv0": ("faster_vit_0_224", {}, dict(
        dim=64, in_dim=64, depths=[2, 3, 6, 5], num_heads=[2, 4, 8, 16], window_size=[7, 7, 7, 7],
        ct_size=2, mlp_ratio=4, resolution=224, hat=[False, False, True, False], do_propagation=False)),
    "fv4": ("faster_vit_4_224", {}, dict(
        dim=196, in_dim=64, depths=[3, 3, 12, 5], num_heads=[4, 8, 16, 32], window_size=[7, 7, 7, 7],
        ct_size=2, mlp_ratio=4, resolution=224, hat=[False, False, True, False], do_propagation=True)),
    "ar0": ("faster_vit_0_any_res", dict(resolution=[576, 960], window_size=[7, 7, 12, 6], ct_size=2, dim=64), dict(
        dim=64, in_dim=64, depths=[2, 3, 6, 5], num_heads=[2, 4, 8, 16], window_size=[7, 7, 12, 6],
        ct_size=2, mlp_ratio=4, resolution=[576, 960], hat=[False, False, True, False],
        do_propagation=False, any_res=True)),
    # reduced models: fv0-like (no layer scale / propagation) and fv4-like (layer scale + propagation,
    # head_dim 12 -> exercises non-power-of-two head dims like fv4's 49)
    "tiny_a": ("faster_vit_0_224", dict(dim=32, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8]), dict(
        dim=32, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 7, 7],
        ct_size=2, mlp_ratio=4, resolution=224, hat=[False, False, True, False], do_propagation=False)),
    "tiny_b": ("faster_vit_4_224", dict(dim=24, in_dim=16, depths=[1, 2, 2, 2], num_heads=[1, 2, 8, 16]), dict(
        dim=24, in_dim=16, depths=[1, 2, 2, 2], num_heads=[1, 2, 8, 16], window_size=[7, 7, 7, 7],
        ct_size=2, mlp_ratio=4, resolution=224, hat=[False, False, True, False], do_propagation=True)),
    # tiny_a with an explicit qk_scale (fv.py:544: scale = qk_scale or head_dim ** -0.5)
    "tiny_qk": ("faster_vit_0_224", dict(dim=32, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], qk_scale=0.3), dict(
        dim=32, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 7, 7],
        ct_size=2, mlp_ratio=4, resolution=224, hat=[False, False, True, False], do_propagation=False, qk_scale=0.3)),
    # any-res reduced: non-square carrier grid (sr = [2, 3]) and padding to the window (30x42 -> 30x42,
    # level 3: 15x21 -> 18x24 with window 6)
    "tiny_ar": ("faster_vit_0_any_res", dict(resolution=[240, 336], window_size=[7, 7, 5, 6], ct_size=2, dim=16,
                                             in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8]), dict(
        dim=16, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 5, 6],
        ct_size=2, mlp_ratio=4, resolution=[240, 336], hat=[False, False, True, False],
        do_propagation=False, any_res=True)),
    # any-res reduced with the larger window sequences the attention kernels special-case: S = 81 + 4 = 85 (one
    # 128-row slot, key loop not needed) and S = 144 + 4 = 148 (BASELINE config 4's level-2 geometry: window 12,
    # ct_size 2, non-square carrier grid) -- forward and backward
    "tiny_ar85": ("faster_vit_0_any_res", dict(resolution=[144, 288], window_size=[7, 7, 9, 5], ct_size=2, dim=16,
                                               in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8]), dict(
        dim=16, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 9, 5],
        ct_size=2, mlp_ratio=4, resolution=[144, 288], hat=[False, False, True, False],
        do_propagation=False, any_res=True)),
    # S = 64 + 4 = 68: one 128-row tile, more than the 64-row window slots of the tile backward -- training takes the
    # key-loop kernels with a single key / query tile (window 8 at level 2, 4 at level 3; level-2 map 16 x 24)
    "tiny_ar68": ("faster_vit_0_any_res", dict(resolution=[256, 384], window_size=[7, 7, 8, 4], ct_size=2, dim=16,
                                               in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8]), dict(
        dim=16, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 8, 4],
        ct_size=2, mlp_ratio=4, resolution=[256, 384], hat=[False, False, True, False],
        do_propagation=False, any_res=True)),
    "tiny_ar148": ("faster_vit_0_any_res", dict(resolution=[384, 576], window_size=[7, 7, 12, 6], ct_size=2, dim=16,
                                                in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8]), dict(
        dim=16, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 12, 6],
        ct_size=2, mlp_ratio=4, resolution=[384, 576], hat=[False, False, True, False],
        do_propagation=False, any_res=True)),
    # 21k fine-tuned family reduced (fv.py:1253-1418): no carrier tokens anywhere (hat all False), layer scale,
    # one 24 x 24 window at level 2 (S = 576: the streaming attention kernel) and 12 x 12 at level 3 (S = 144)
    "tiny_21k": ("faster_vit_4_21k_384", dict(dim=24, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8]), dict(
        dim=24, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 24, 12],
        ct_size=2, mlp_ratio=4, resolution=384, hat=[False, False, False, False], do_propagation=True)),
    # faster_vit_4_21k_224 reduced (fv.py:1253-1290): one 14 x 14 window at level 2 (S = 196: the key-loop tensor-core
    # attention forward AND backward, two 128-row tiles), 7 x 7 at level 3 -- the member of the 21k family that trains
    "tiny_21k224": ("faster_vit_4_21k_224", dict(dim=24, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8]), dict(
        dim=24, in_dim=16, depths=[1, 1, 2, 1], num_heads=[1, 2, 4, 8], window_size=[7, 7, 14, 7],
        ct_size=2, mlp_ratio=4, resolution=224, hat=[False, False, False, False], do_propagation=True)),
}


def cfg_of(name: str) -> dict:
    return dict(CASES[name][2])
