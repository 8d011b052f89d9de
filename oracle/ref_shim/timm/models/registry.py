_entrypoints = {}


def register_model(fn):
    """timm.models.registry.register_model: record the entrypoint under its function name."""
    _entrypoints[fn.__name__] = fn
    return fn
