import importlib
import sys


class _Cfg:
    def __init__(self, d):
        self._d = dict(d)

    def to_dict(self):
        return dict(self._d)


def resolve_pretrained_cfg(name, **_):
    for modname in ("fastervit.models.faster_vit", "fastervit.models.faster_vit_any_res"):
        mod = sys.modules.get(modname)
        if mod is not None and name in getattr(mod, "default_cfgs", {}):
            return _Cfg(mod.default_cfgs[name])
    raise KeyError(name)


def _update_default_model_kwargs(pretrained_cfg, kwargs, kwargs_filter=None):
    kwargs.setdefault("num_classes", pretrained_cfg["num_classes"])
    kwargs.setdefault("in_chans", pretrained_cfg["input_size"][0])
