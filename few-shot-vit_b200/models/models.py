"""Model registry -- the plugin API the reference's scripts call (reference: test_phase/models/models.py:4-27;
`strict=False` loading as in sun_meta_training/models/models.py:25 is available through `load(..., strict=False)`)."""
import torch

models = {}


def register(name):
    """Decorator: `@register('meta-baseline')` stores the class / factory under `name`."""
    def add(obj):
        models[name] = obj
        return obj
    return add


def make(name, **kwargs):
    """Instantiate a registered model; like the reference it lands on the GPU when one is present."""
    if name is None:
        return None
    if name not in models:
        raise KeyError(f"model '{name}' is not registered (known: {sorted(models)})")
    net = models[name](**kwargs)
    return net.cuda() if torch.cuda.is_available() else net


def load(model_sv, name=None, strict=True):
    """Rebuild a model from a checkpoint dict {name, name_args, name_sd} (reference save format, train_meta.py:241-266)."""
    key = name or "model"
    net = make(model_sv[key], **model_sv[key + "_args"])
    net.load_state_dict(model_sv[key + "_sd"], strict=strict)
    return net
