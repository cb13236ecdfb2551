"""Checkpoint loading for the drop-in models — the host logic on the caller side of the path that turns a released /
DAPT checkpoint into the classifier's state dict (run_frame_finetuning.py:396-460, cited as rff; utils.py:336-383, ut).

    select_state_dict      rff:404-411   pick checkpoint["model"] / ["module"] (args.model_key = "model|module")
    remap_finetune_keys    rff:413-430   drop a head of another width; backbone.* -> *, encoder.norm.* -> fc_norm.*,
                                         encoder.* -> *   (a DAPT / VideoMAE pre-training checkpoint feeds the classifier)
    interpolate_pos_embed  rff:432-458   bicubic resize of a learnable position table to the model's patch grid
    load_state_dict        ut:336-383    non-strict load that reports, rather than raises on, missing / unexpected keys
    load_finetune_checkpoint             the four in the order of rff:396-460

Pure host code (torch on the CPU, once per model); nothing here is on the forward path.
"""
from collections import OrderedDict

import torch


def select_state_dict(checkpoint, model_key="model|module"):
    """rff:404-411: the first of the `|`-separated keys present in the checkpoint, else the checkpoint itself."""
    for key in model_key.split('|'):
        if key in checkpoint:
            return checkpoint[key]
    return checkpoint


def remap_finetune_keys(checkpoint_model, model):
    """rff:413-430.  Returns a new OrderedDict; `checkpoint_model` loses a mismatched head (as in the reference)."""
    own = model.state_dict()
    for k in ('head.weight', 'head.bias'):
        if k in checkpoint_model and k in own and checkpoint_model[k].shape != own[k].shape:
            del checkpoint_model[k]                                   # rff:414-417 (classifier of another width)
    out = OrderedDict()
    for key in list(checkpoint_model.keys()):
        if key.startswith('backbone.'):
            out[key[9:]] = checkpoint_model[key]
        elif key.startswith('encoder.norm'):
            out[key.replace("encoder.norm", "fc_norm")] = checkpoint_model[key]
        elif key.startswith('encoder.'):
            out[key[8:]] = checkpoint_model[key]
        else:
            out[key] = checkpoint_model[key]
    return out


def interpolate_pos_embed(checkpoint_model, model, num_frames=16):
    """rff:432-458: if the checkpoint carries a (learnable) `pos_embed` whose spatial grid differs from the model's, the
    position rows are resized bicubically per temporal slot; leading extra tokens are kept.  In place; returns the dict."""
    if 'pos_embed' not in checkpoint_model:
        return checkpoint_model
    pos = checkpoint_model['pos_embed']
    C = pos.shape[-1]
    num_patches = model.patch_embed.num_patches
    n_extra = model.pos_embed.shape[-2] - num_patches                 # 0 / 1
    t_slots = num_frames // model.patch_embed.tubelet_size
    orig = int(((pos.shape[-2] - n_extra) // t_slots) ** 0.5)
    new = int((num_patches // t_slots) ** 0.5)
    if orig != new:
        extra = pos[:, :n_extra]
        tok = pos[:, n_extra:].reshape(-1, t_slots, orig, orig, C).reshape(-1, orig, orig, C).permute(0, 3, 1, 2)
        tok = torch.nn.functional.interpolate(tok, size=(new, new), mode='bicubic', align_corners=False)
        tok = tok.permute(0, 2, 3, 1).reshape(-1, t_slots, new, new, C).flatten(1, 3)
        checkpoint_model['pos_embed'] = torch.cat((extra, tok), dim=1)
    return checkpoint_model


def load_state_dict(model, state_dict, prefix='', ignore_missing="relative_position_index", verbose=True):
    """ut:336-383: load what matches, report the rest.  Keys are looked up as `prefix + name`.  Returns the missing keys
    (those not matched by one of the `|`-separated `ignore_missing` fragments).  A tensor of the wrong shape is reported
    and skipped (the reference collects it in error_msgs and prints it), never raised."""
    own = model.state_dict()
    usable, errors, unexpected = OrderedDict(), [], []
    for key, value in state_dict.items():
        if prefix and not key.startswith(prefix):
            unexpected.append(key)
            continue
        name = key[len(prefix):]
        if name not in own:
            unexpected.append(key)
        elif tuple(own[name].shape) != tuple(value.shape):
            errors.append(f"size mismatch for {key}: copying a param with shape {tuple(value.shape)} from checkpoint, "
                          f"the shape in current model is {tuple(own[name].shape)}.")
        else:
            usable[name] = value
    model.load_state_dict(usable, strict=False)
    missing_all = [k for k in own if k not in usable]
    fragments = [f for f in ignore_missing.split('|') if f]
    missing = [k for k in missing_all if not any(f in k for f in fragments)]
    ignored = [k for k in missing_all if k not in missing]
    if verbose:
        name = model.__class__.__name__
        if missing:
            print("Weights of {} not initialized from pretrained model: {}".format(name, missing))
        if unexpected:
            print("Weights from pretrained model not used in {}: {}".format(name, unexpected))
        if ignored:
            print("Ignored weights of {} not initialized from pretrained model: {}".format(name, ignored))
        if errors:
            print('\n'.join(errors))
    return missing


def load_finetune_checkpoint(model, checkpoint, model_key="model|module", model_prefix='', num_frames=16, verbose=True):
    """rff:396-460 for a checkpoint path or an already loaded dict.  Returns the missing keys."""
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__"):
        checkpoint = torch.load(checkpoint, map_location='cpu')
    checkpoint_model = select_state_dict(checkpoint, model_key)
    checkpoint_model = remap_finetune_keys(checkpoint_model, model)
    checkpoint_model = interpolate_pos_embed(checkpoint_model, model, num_frames)
    return load_state_dict(model, checkpoint_model, prefix=model_prefix, verbose=verbose)
