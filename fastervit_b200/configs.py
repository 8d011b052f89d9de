"""Entrypoint table of the 22 FasterViT variants (11 square + 11 any-resolution).

Values restate the keyword defaults of the reference entrypoints (fastervit/models/faster_vit.py:977-1418,
faster_vit_any_res.py:1007-1448) and their `default_cfgs` (faster_vit.py:21-80): same names, same
overridable kwargs, same pretrained_cfg fields, so `create_model(name, **kwargs)` is a drop-in.
"""
from __future__ import annotations

import copy

_HF = "https://huggingface.co/ahatamiz/FasterViT/resolve/main/"

# (suffix, depths, heads, dim, in_dim, drop_path, layer_scale, do_propagation, ckpt stem)
_FAMILY = {
    "0": ([2, 3, 6, 5], [2, 4, 8, 16], 64, 64, 0.2, None, False, "faster_vit_0"),
    "1": ([1, 3, 8, 5], [2, 4, 8, 16], 80, 32, 0.2, None, False, "faster_vit_1"),
    "2": ([3, 3, 8, 5], [2, 4, 8, 16], 96, 64, 0.2, None, False, "faster_vit_2"),
    "3": ([3, 3, 12, 5], [2, 4, 8, 16], 128, 64, 0.3, 1e-5, True, "faster_vit_3"),
    "4": ([3, 3, 12, 5], [4, 8, 16, 32], 196, 64, 0.3, 1e-5, True, "faster_vit_4"),
    "5": ([3, 3, 12, 5], [4, 8, 16, 32], 320, 64, 0.3, 1e-5, True, "faster_vit_5"),
    "6": ([3, 3, 16, 8], [4, 8, 16, 32], 320, 64, 0.5, 1e-5, True, "faster_vit_6"),
}
# ImageNet-21k fine-tuned FasterViT-4 at larger windows: (resolution, window sizes, crop_pct)
_21K = {"224": (224, [7, 7, 14, 7], 0.95), "384": (384, [7, 7, 24, 12], 1.0),
        "512": (512, [7, 7, 32, 16], 1.0), "768": (768, [7, 7, 48, 24], 0.93)}
_CROP = {"0": 0.875}


def _pcfg(url: str, input_size, crop_pct: float, crop_mode: str) -> dict:
    return {"url": url, "num_classes": 1000, "input_size": tuple(input_size), "pool_size": None,
            "crop_pct": crop_pct, "interpolation": "bicubic", "fixed_input_size": True,
            "mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225), "crop_mode": crop_mode}


def _build() -> tuple[dict, dict]:
    models, cfgs = {}, {}
    for sfx, (depths, heads, dim, in_dim, dpr, ls, prop, stem) in _FAMILY.items():
        base = dict(depths=depths, num_heads=heads, window_size=[7, 7, 7, 7], ct_size=2, dim=dim,
                    in_dim=in_dim, mlp_ratio=4, drop_path_rate=dpr, hat=[False, False, True, False],
                    model_path=f"/tmp/{stem}.pth.tar")
        fixed = dict(do_propagation=prop)
        if ls is not None:
            base["layer_scale"] = ls
            fixed["layer_scale_conv"] = None
        pc = _pcfg(_HF + f"fastervit_{sfx}_224_1k.pth.tar", (3, 224, 224), _CROP.get(sfx, 1.0), "center")
        models[f"faster_vit_{sfx}_224"] = dict(defaults=dict(base, resolution=224), fixed=fixed, any_res=False)
        cfgs[f"faster_vit_{sfx}_224"] = pc
        ar_res = [541, 960] if sfx == "2" else [576, 960]
        models[f"faster_vit_{sfx}_any_res"] = dict(defaults=dict(base, resolution=ar_res), fixed=fixed,
                                                   any_res=True)
        cfgs[f"faster_vit_{sfx}_any_res"] = copy.deepcopy(pc)
    depths, heads, dim, in_dim, _, ls, prop, _ = _FAMILY["4"]
    for res_name, (res, ws, crop) in _21K.items():
        base = dict(depths=depths, num_heads=heads, window_size=ws, ct_size=2, dim=dim, in_dim=in_dim,
                    mlp_ratio=4, layer_scale=ls, hat=[False, False, False, False],
                    model_path=f"/tmp/fastervit_4_21k_{res_name}_w{ws[2]}.pth.tar")
        fixed = dict(do_propagation=prop, layer_scale_conv=None)
        pc = _pcfg(_HF + f"fastervit_4_21k_{res_name}_w{ws[2]}.pth.tar", (3, res, res), crop, "squash")
        models[f"faster_vit_4_21k_{res_name}"] = dict(
            defaults=dict(base, resolution=res, drop_path_rate=0.42), fixed=fixed, any_res=False)
        cfgs[f"faster_vit_4_21k_{res_name}"] = pc
        models[f"faster_vit_4_21k_{res_name}_any_res"] = dict(
            defaults=dict(base, resolution=[576, 960], drop_path_rate=0.3), fixed=fixed, any_res=True)
        cfgs[f"faster_vit_4_21k_{res_name}_any_res"] = copy.deepcopy(pc)
    return models, cfgs


MODEL_SPECS, default_cfgs = _build()
