"""B200-native drop-in for the reference's other_models/UMT/modeling_finetune.py (cited below as umt:line).

UMT fine-tuning uses the simple-tad Video-ViT unchanged except for the position table (umt:195-239, umt:282-293): the
1-D sinusoid table of the 8-frame 14x14 pre-training grid (1568 rows; 2048 for patch 14) is interpolated — bicubic over
(h', w') when the image grid differs, linear over t' when the number of temporal slots differs — and becomes a learnable
parameter whenever the result has a different number of rows.  The reference's job (other_models/UMT/u_dada_base_a100.sh)
runs it with tubelet_size = 1 on 8 frames, i.e. 1568 tokens and the plain table.  The interpolation is host-side model
construction (torch on the CPU, once per model); the forward is one stad_vit_forward call.
"""
from functools import partial

import torch
import torch.nn as nn

from ...modeling_finetune import (Attention, Block, DropPath, Mlp, PatchEmbed,  # noqa: F401
                                  VisionTransformer as _VisionTransformer, _cfg, get_sinusoid_encoding_table as _sinusoid)
from ...registry import register_model

__all__ = [
    "Mlp", "Attention", "Block", "PatchEmbed", "DropPath", "get_sinusoid_encoding_table", "VisionTransformer",
    "vit_small_patch16_224", "vit_base_patch16_224", "vit_base_patch16_384", "vit_large_patch16_224",
    "vit_large_patch16_384", "vit_large_patch16_512", "vit_huge_patch16_224",
]

_CKPT_T, _CKPT_P = 8, 14  # temporal slots and grid side of the pre-training checkpoint (umt:211-212, umt:225)


def get_sinusoid_encoding_table(n_position, d_hid, cur_frame=-1, pre_n_position=1568):
    """[1, n_position, d_hid] position table (umt:195-239): the sinusoid table over `pre_n_position` flat indices,
    resampled to the current grid.  A tensor when n_position == pre_n_position, else an nn.Parameter."""
    table = _sinusoid(pre_n_position, d_hid)                                  # umt:201-205
    C = d_hid
    if cur_frame != -1 and n_position // cur_frame * 8 != pre_n_position:      # umt:208-221: other image grid
        T, P = _CKPT_T, _CKPT_P
        new_P = int((n_position // cur_frame) ** 0.5)
        grid = table.reshape(-1, T, P, P, C).reshape(-1, P, P, C).permute(0, 3, 1, 2)
        grid = torch.nn.functional.interpolate(grid, size=(new_P, new_P), mode='bicubic', align_corners=False)
        table = grid.permute(0, 2, 3, 1).reshape(-1, T, new_P, new_P, C).flatten(1, 3)
    if cur_frame != -1 and cur_frame != 8:                                     # umt:222-234: other clip length
        T, new_T = _CKPT_T, cur_frame
        P = int((n_position // cur_frame) ** 0.5)
        line = table.reshape(-1, T, P, P, C).permute(0, 2, 3, 4, 1).reshape(-1, C, T)
        line = torch.nn.functional.interpolate(line, size=new_T, mode='linear')
        table = line.reshape(1, P, P, C, new_T).permute(0, 4, 1, 2, 3).flatten(1, 3)
    if n_position == pre_n_position:
        return table
    return nn.Parameter(table, requires_grad=True)                             # umt:237-239


class VisionTransformer(_VisionTransformer):
    """UMT VisionTransformer (umt:242-373): the simple-tad model with the interpolated position table."""

    def _build_pos_embed(self, num_patches, embed_dim, learnable):
        if learnable:                                                          # umt:282-283
            return nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        pe = self.patch_embed
        pre_n_position = 2048 if pe.patch_size[0] == 14 else 1568              # umt:286-289
        return get_sinusoid_encoding_table(num_patches, embed_dim, pe.num_frames // pe.tubelet_size,
                                           pre_n_position=pre_n_position)      # umt:290-293


_SPECS = {
    "vit_small_patch16_224": (224, 384, 12, 6),
    "vit_base_patch16_224": (224, 768, 12, 12),
    "vit_base_patch16_384": (384, 768, 12, 12),
    "vit_large_patch16_224": (224, 1024, 24, 16),
    "vit_large_patch16_384": (384, 1024, 24, 16),
    "vit_large_patch16_512": (512, 1024, 24, 16),
    "vit_huge_patch16_224": (224, 1280, 32, 16),
}


def _factory(name, img_size, embed_dim, depth, num_heads):
    def make(pretrained=False, **kwargs):
        kwargs.setdefault("img_size", img_size)
        model = VisionTransformer(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4,
                                  qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
        model.default_cfg = _cfg()
        return model
    make.__name__ = make.__qualname__ = name
    make.__doc__ = f"UMT {name}: img {img_size}, D={embed_dim}, depth={depth}, heads={num_heads} (umt:375-436)."
    return make


# Same factory names as simple_tad_b200.modeling_finetune (the reference's UMT runner imports its own module and the
# names shadow each other in timm's registry, umt:375): registered under a "umt_" prefix in this package's registry,
# exported under the reference's names from this module.
for _name, _spec in _SPECS.items():
    _fn = _factory(_name, *_spec)
    globals()[_name] = _fn
    _alias = _factory("umt_" + _name, *_spec)
    register_model(_alias)
del _name, _spec, _fn, _alias
