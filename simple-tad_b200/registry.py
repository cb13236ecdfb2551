"""Minimal stand-in for the two timm entry points the reference uses for this path:
`timm.models.registry.register_model` (modeling_finetune.py:7, modeling_pretrain.py:9) and
`timm.models.create_model` (run_frame_finetuning.py:374-389, run_inference.py:40-55, test_efficiency.py:24-57).

timm 0.4.12's create_model drops None-valued kwargs before calling the factory (the callers pass
`drop_block_rate=None`, which the VisionTransformer ctor does not accept); that filtering is reproduced.  If timm is
importable the factories are registered there too, so `timm.models.create_model(name)` resolves to this package."""
_model_entrypoints = {}


def register_model(fn):
    _model_entrypoints[fn.__name__] = fn
    try:  # pragma: no cover - timm is not installed in the build image
        from timm.models.registry import register_model as _timm_register
        _timm_register(fn)
    except Exception:
        pass
    return fn


def list_models():
    return sorted(_model_entrypoints)


def is_model(name):
    return name in _model_entrypoints


def create_model(model_name, pretrained=False, checkpoint_path="", **kwargs):
    if model_name not in _model_entrypoints:
        raise RuntimeError("Unknown model (%s)" % model_name)
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    model = _model_entrypoints[model_name](pretrained=pretrained, **kwargs)
    if checkpoint_path:
        import torch
        state = torch.load(checkpoint_path, map_location="cpu")
        for key in ("model", "module", "state_dict"):
            if isinstance(state, dict) and key in state:
                state = state[key]
                break
        model.load_state_dict(state)
    return model
