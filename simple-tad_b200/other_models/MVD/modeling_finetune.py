"""B200-native drop-in for the reference's other_models/MVD/modeling_finetune.py (cited below as mvd:line).

The MVD student is the simple-tad Video-ViT with two differences (mvd:322-458):
  * the position table is a fixed 3-D sin-cos embedding — D/4 channels encode t', 3D/4 channels the (h', w') grid —
    instead of the 1-D sinusoid over the flat token index (mvd:24-69, mvd:370-373);
  * an optional class token (`use_cls_token`, mvd:361-368) is prepended AFTER the position add (mvd:428-435), so it
    carries no position row; 'fc_norm' pooling skips it (mvd:447-449); every other `final_reduction` returns
    norm(x)[:, 0] (mvd:450-451 — MVD has no per-token 'none' output).
Blocks, attention, MLP and the patch embedding are those of simple_tad_b200.modeling_finetune; the forward is one
stad_vit_forward call (the class token is laid out by stad_prepend_cls inside it).
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from ... import _lib
from ...modeling_finetune import (Attention, Block, DropPath, Mlp, PatchEmbed as _PatchEmbed,  # noqa: F401
                                  VisionTransformer as _VisionTransformer, _cfg, get_sinusoid_encoding_table as _sinusoid,
                                  trunc_normal_)
from ...registry import register_model

__all__ = [
    "Mlp", "Attention", "Block", "PatchEmbed", "DropPath", "get_3d_sincos_pos_embed", "get_2d_sincos_pos_embed",
    "get_2d_sincos_pos_embed_from_grid", "get_1d_sincos_pos_embed_from_grid", "get_sinusoid_encoding_table",
    "VisionTransformer", "mvd_vit_small_patch16_224", "mvd_vit_base_patch16_224", "mvd_vit_large_patch16_224",
    "mvd_vit_huge_patch16_224",
]


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos, scale=None):
    """[M, embed_dim] = (sin(pos x omega) | cos(pos x omega)), omega_k = 10000^(-k / (embed_dim / 2)) in float64
    (mvd:102-122)."""
    assert embed_dim % 2 == 0
    half = embed_dim // 2
    omega = 1.0 / 10000 ** (np.arange(half, dtype=float) / (embed_dim / 2.0))
    pos = np.asarray(pos).reshape(-1)
    if scale is not None:
        pos = pos * scale
    angle = np.einsum("m,d->md", pos, omega)
    return np.concatenate([np.sin(angle), np.cos(angle)], axis=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    """First half of the channels from grid[0], second half from grid[1] (mvd:89-100)."""
    assert embed_dim % 2 == 0
    return np.concatenate([get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0]),
                           get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])], axis=1)


def _grid_wh(grid_size):
    # np.meshgrid(w, h): plane 0 holds the w' coordinate, plane 1 the h' coordinate, both laid out [h', w'] (mvd:38-43)
    axis = np.arange(grid_size, dtype=np.float32)
    return np.stack(np.meshgrid(axis, axis), axis=0).reshape(2, 1, grid_size, grid_size)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    """[grid_size^2 (+1), embed_dim] numpy table (mvd:72-87)."""
    table = get_2d_sincos_pos_embed_from_grid(embed_dim, _grid_wh(grid_size))
    if cls_token:
        table = np.concatenate([np.zeros([1, embed_dim]), table], axis=0)
    return table


def get_3d_sincos_pos_embed(embed_dim, grid_size, t_size, cls_token=False, scale_t=None):
    """[1, t_size * grid_size^2 (+1), embed_dim] fp32: channels [0, D/4) encode t', channels [D/4, D) the 2-D grid; token
    order (t', h', w') (mvd:24-69)."""
    assert embed_dim % 4 == 0
    d_t = embed_dim // 4
    d_s = embed_dim // 4 * 3
    spatial = get_2d_sincos_pos_embed_from_grid(d_s, _grid_wh(grid_size))                       # [H*W, 3D/4]
    temporal = get_1d_sincos_pos_embed_from_grid(d_t, np.arange(t_size, dtype=np.float32), scale=scale_t)  # [T, D/4]
    hw = grid_size ** 2
    table = np.concatenate([np.repeat(temporal[:, None, :], hw, axis=1),
                            np.repeat(spatial[None, :, :], t_size, axis=0)], axis=-1).reshape(-1, embed_dim)
    if cls_token:
        table = np.concatenate([np.zeros([1, embed_dim]), table], axis=0)
    return torch.FloatTensor(table).unsqueeze(0)


def get_sinusoid_encoding_table(n_position, d_hid, cls_token=False):
    """The 1-D sinusoid table; MVD's copy accepts and ignores `cls_token` (mvd:310-320)."""
    return _sinusoid(n_position, d_hid)


class PatchEmbed(_PatchEmbed):
    """PatchEmbed with the grid attributes MVD adds (mvd:290-292)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2):
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                         num_frames=num_frames, tubelet_size=tubelet_size)
        self.num_patches_w = self.img_size[1] // self.patch_size[1]
        self.num_patches_h = self.img_size[0] // self.patch_size[0]
        self.num_patches_t = num_frames // self.tubelet_size


class VisionTransformer(_VisionTransformer):
    """MVD VisionTransformer (mvd:322-458), sm_100a forward."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, fc_drop_rate=0., drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=0., use_flash_attn=True, init_scale=0.,
                 all_frames=16, tubelet_size=2, use_checkpoint=False, final_reduction="fc_norm", use_cls_token=False):
        self.use_cls_token = use_cls_token
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes,
                         embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                         qk_scale=qk_scale, fc_drop_rate=fc_drop_rate, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
                         drop_path_rate=drop_path_rate, norm_layer=norm_layer, init_values=init_values,
                         use_learnable_pos_emb=False, use_flash_attn=use_flash_attn, init_scale=init_scale,
                         all_frames=all_frames, tubelet_size=tubelet_size, use_checkpoint=use_checkpoint,
                         final_reduction=final_reduction)
        self.patch_size = patch_size
        pe = self.patch_embed
        pe.num_patches_w = pe.img_size[1] // pe.patch_size[1]
        pe.num_patches_h = pe.img_size[0] // pe.patch_size[0]
        pe.num_patches_t = pe.num_frames // pe.tubelet_size

    def _build_pos_embed(self, num_patches, embed_dim, learnable):
        pe = self.patch_embed
        return get_3d_sincos_pos_embed(embed_dim=embed_dim, grid_size=pe.img_size[0] // pe.patch_size[0],
                                       t_size=pe.num_frames // pe.tubelet_size)     # mvd:370-373

    def _init_extra_tokens(self):
        if self.use_cls_token:                                                       # mvd:364-368, mvd:391-392
            self.cls_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
            trunc_normal_(self.cls_token, std=.02)
        else:
            self.cls_token = None

    def _cls_token(self):
        return self.cls_token

    def _reduction(self):
        # mvd:445-451: fc_norm pools the patch tokens; every other mode returns norm(x)[:, 0]
        if self.final_reduction == "fc_norm":
            return _lib.STAD_REDUCE_MEAN, self.fc_norm
        return _lib.STAD_REDUCE_CLS, self.norm



def _factory(name, embed_dim, depth, num_heads):
    def make(pretrained=False, **kwargs):
        model = VisionTransformer(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4,
                                  qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
        model.default_cfg = _cfg()
        return model
    make.__name__ = make.__qualname__ = name
    make.__doc__ = f"{name}: D={embed_dim}, depth={depth}, heads={num_heads} (mvd:459-492)."
    return register_model(make)


mvd_vit_small_patch16_224 = _factory("mvd_vit_small_patch16_224", 384, 12, 6)
mvd_vit_base_patch16_224 = _factory("mvd_vit_base_patch16_224", 768, 12, 12)
mvd_vit_large_patch16_224 = _factory("mvd_vit_large_patch16_224", 1024, 24, 16)
mvd_vit_huge_patch16_224 = _factory("mvd_vit_huge_patch16_224", 1280, 32, 16)
