"""B200-native drop-in for the visible-token encoder of the reference's modeling_pretrain.py (DAPT / MAE pre-training).

In scope: `PretrainVisionTransformerEncoder` (modeling_pretrain.py:26-113) — embed, add the position
table, keep the visible tokens, run the blocks, apply `norm`.  Unlike the reference, which embeds all 1568 tokens and
then throws 90 % of them away (mp:93-98), only the visible tokens are embedded.  The MAE decoder
(modeling_pretrain.py:115-291) is the next row of the scope table (§8f) and is not built yet: the
`pretrain_videomae_*` factories return the encoder.
"""
from functools import partial

import torch
import torch.nn as nn

from . import _lib
from .modeling_finetune import Block, PatchEmbed, _StadBackbone, _cfg, _inference_only, get_sinusoid_encoding_table
from .registry import register_model

__all__ = [
    "PretrainVisionTransformerEncoder",
    "pretrain_videomae_small_patch16_224",
    "pretrain_videomae_base_patch16_224",
    "pretrain_videomae_large_patch16_224",
    "pretrain_videomae_huge_patch16_224",
]


def visible_token_indices(mask, n_visible=None):
    """Row-major ids of the tokens x[~mask] keeps (mp:98), as int32 [B, n_visible], computed on the device.
    Every clip must keep the same number of tokens (tube masking guarantees it, masking_generator.py:8-9)."""
    if mask.dim() != 2:
        raise ValueError(f"mask must be [B, N] bool, got {tuple(mask.shape)}")
    mask = mask.bool()
    if n_visible is None:
        n_visible = int((~mask[0]).sum())  # one host sync, as x[~mask] has in the reference
    # stable argsort of the mask puts the visible (False) positions first, in their original order
    order = torch.argsort(mask.to(torch.uint8), dim=1, stable=True)
    return order[:, :n_visible].to(torch.int32).contiguous(), n_visible


class PretrainVisionTransformerEncoder(_StadBackbone):
    """MAE encoder over the visible tokens (modeling_pretrain.py:26-113), sm_100a forward."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=None, tubelet_size=2, use_checkpoint=False,
                 use_learnable_pos_emb=False, use_flash_attn=True):
        super().__init__()
        self.num_classes = num_classes
        self.num_heads = num_heads
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(
            img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, tubelet_size=tubelet_size)
        num_patches = self.patch_embed.num_patches
        self.use_checkpoint = use_checkpoint
        if use_learnable_pos_emb:
            raise NotImplementedError("use_learnable_pos_emb allocates num_patches + 1 rows in the reference (mp:47) "
                                      "and no script enables it")
        self.pos_embed = get_sinusoid_encoding_table(num_patches, embed_dim)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(
                dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                init_values=init_values, use_flash_attn=use_flash_attn)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_num_layers(self):
        return len(self.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def prepare(self, device=None):
        _inference_only(self)
        if not isinstance(self.head, nn.Identity):
            raise NotImplementedError("encoder_num_classes > 0 is never used by the reference (mp:203); head must be Identity")
        self.blocks[0].attn._check_head_dim()
        device = device or next(self.parameters()).device
        return self._prepared_for(torch.device(device), self.norm, None)

    @torch.no_grad()
    def forward_features(self, x, mask, n_visible=None):
        """x [B, C, T, H, W], mask [B, N] bool (True = masked) -> [B, N_vis, D] fp32 after `norm` (mp:91-108)."""
        if not x.is_cuda:
            raise RuntimeError("simple-tad_b200 runs on a CUDA (sm_100a) device only; got a CPU tensor")
        B, Cc, T, H, W = x.shape
        pe = self.patch_embed
        assert H == pe.img_size[0] and W == pe.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]})."
        if mask.shape != (B, pe.num_patches):
            raise ValueError(f"mask must be [{B}, {pe.num_patches}], got {tuple(mask.shape)}")
        prep = self.prepare(x.device)
        tok_idx, n_vis = visible_token_indices(mask.to(x.device), n_visible)
        xb = prep.input_bf16(x)
        inp = _lib.make_input(xb, _lib.STAD_IN_CLIPS)
        return prep.run(inp, B, n_vis, tok_idx=tok_idx, want=("tokens",))["tokens"]

    @torch.no_grad()
    def forward(self, x, mask, n_visible=None):
        return self.forward_features(x, mask, n_visible)  # head = Identity (mp:110-113)


_ENCODER_SPECS = {
    # reference factories modeling_pretrain.py:293-387 (encoder part): name -> (embed_dim, depth, num_heads)
    "pretrain_videomae_small_patch16_224": (384, 12, 6),
    "pretrain_videomae_base_patch16_224": (768, 12, 12),
    "pretrain_videomae_large_patch16_224": (1024, 24, 16),
    "pretrain_videomae_huge_patch16_224": (1280, 32, 16),
}
_DECODER_ONLY_KWARGS = ("decoder_depth", "decoder_embed_dim", "decoder_num_heads", "decoder_num_classes",
                        "encoder_num_classes", "encoder_in_chans")


def _make_encoder_factory(name, embed_dim, depth, num_heads):
    def factory(pretrained=False, **kwargs):
        for k in _DECODER_ONLY_KWARGS:  # accepted for call-site parity; the decoder is not part of this build yet
            kwargs.pop(k, None)
        model = PretrainVisionTransformerEncoder(
            img_size=224, patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, num_classes=0,
            mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
        model.default_cfg = _cfg()
        return model
    factory.__name__ = factory.__qualname__ = name
    factory.__doc__ = f"{name}: visible-token encoder D={embed_dim}, depth={depth}, heads={num_heads}."
    return register_model(factory)


for _name, _spec in _ENCODER_SPECS.items():
    globals()[_name] = _make_encoder_factory(_name, *_spec)
del _name, _spec
