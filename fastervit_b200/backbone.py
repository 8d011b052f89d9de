"""Dense-prediction backbone interface of FasterViT — the surface the reference's downstream copies
expose (downstream/object_detection/dino/models/dino/fastervit.py:686-846): per-level feature maps taken *before* each
level's Downsample (`xo`, fastervit.py:705-709), one BatchNorm2d `norm{i}` per requested level (fastervit.py:795-799),
`forward_raw(x) -> tuple of [B, C_i, H_i, W_i]`, `forward(tensors, mask) -> {idx: (features, mask)}` with the padding
mask resized to every level (fastervit.py:829-846), frozen stages (fastervit.py:803-815).

The hot path is the same sm_100a launch list as classification: the wrapped model's eval plan is run up to the points
where the level outputs are live and `fvit_feature_map_fwd` applies the level's folded BatchNorm while converting the
token-major activation to NCHW. Variable input sizes go through the any-res entrypoints (one cached plan per shape; the
level geometry is padded to window multiples exactly as faster_vit_any_res.py:851-867 does).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as L
from .registry import create_model


class FasterViTBackbone(nn.Module):
    def __init__(self, model_name: str = "faster_vit_0_any_res", out_indices=(0, 1, 2, 3), frozen_stages: int = -1,
                 norm_layer=nn.BatchNorm2d, **kwargs):
        super().__init__()
        kwargs.setdefault("num_classes", 0)
        self.body = create_model(model_name, **kwargs)
        dim = self.body.cfg["dim"]
        self.num_levels = len(self.body.levels)
        self.num_features = [int(dim * 2 ** i) for i in range(self.num_levels)]
        self.out_indices = tuple(out_indices)
        if norm_layer is not nn.BatchNorm2d:
            raise NotImplementedError("the per-level norm is the reference default (BatchNorm2d), folded into the "
                                      "layout-conversion kernel")
        for i in self.out_indices:
            self.add_module(f"norm{i}", nn.BatchNorm2d(self.num_features[i]))
        self.frozen_stages = frozen_stages
        self._freeze_stages()

    def _freeze_stages(self):   # fastervit.py:803-815
        if self.frozen_stages >= 0:
            self.body.patch_embed.eval()
            for p in self.body.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 2:
            for i in range(0, self.frozen_stages - 1):
                m = self.body.levels[i]
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def train(self, mode: bool = True):
        super().train(mode)
        self._freeze_stages()
        return self

    @torch.no_grad()
    def forward_raw(self, x: torch.Tensor) -> tuple:
        """fastervit.py:816-827: tuple of the normalised level outputs, NCHW."""
        if self.training:
            raise L.FvitError("FasterViTBackbone runs the inference launch list (call .eval()); fine-tuning a detector "
                              "end to end through the backbone is outside the classification hot path")
        feats = self.body.forward_levels(x, self.out_indices)
        outs = []
        for idx, f in zip(self.out_indices, feats):
            bn: nn.BatchNorm2d = getattr(self, f"norm{idx}")
            B, C, H, W = f.shape
            scale = (bn.weight * torch.rsqrt(bn.running_var + bn.eps)).float().contiguous()
            shift = (bn.bias - bn.running_mean * scale).float().contiguous()
            rows = f.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()     # token-major view of the level output
            out = torch.empty(B, C, H, W, dtype=torch.float32, device=f.device)
            with torch.cuda.device(f.device):
                L.call("fvit_feature_map_fwd", rows.data_ptr(), C, None, B, H * W, C, scale.data_ptr(), shift.data_ptr(),
                       out.data_ptr())
            outs.append(out)
        return tuple(outs)

    @torch.no_grad()
    def forward(self, tensors: torch.Tensor, mask: torch.Tensor | None = None) -> dict:
        """fastervit.py:829-846 without the NestedTensor wrapper type: {idx: (features, mask resized to the level)}."""
        outs = self.forward_raw(tensors)
        res = {}
        for i, o in enumerate(outs):
            m = None
            if mask is not None:
                m = torch.nn.functional.interpolate(mask[None].float(), size=o.shape[-2:]).to(torch.bool)[0]
            res[i] = (o, m)
        return res
