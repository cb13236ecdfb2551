"""B200-native drop-in for the reference's modeling_pretrain.py (DAPT / MAE pre-training forward).

SURVEY §8 a11: `PretrainVisionTransformerEncoder` (modeling_pretrain.py:26-113) — embed, add the position table, keep
the visible tokens, run the blocks, apply `norm`.  Unlike the reference, which embeds all 1568 tokens and then throws
90 % of them away (mp:93-98), only the visible tokens are embedded.
SURVEY §8 f1: `PretrainVisionTransformerDecoder` (mp:115-180) and the full `PretrainVisionTransformer` (mp:183-291):
encoder -> encoder_to_decoder -> position rows + mask tokens -> decoder blocks -> norm -> pixel head on the masked
tokens, one `stad_mae_forward` call.  The `pretrain_videomae_*` factories return the full model, as in the reference.
"""
import ctypes as C
from functools import partial

import torch
import torch.nn as nn

from . import _lib
from .modeling_finetune import (Block, PatchEmbed, _StadBackbone, _PreparedModel, _as_bf16_2d, _cfg, _inference_only,
                                _weights_signature, get_sinusoid_encoding_table, trunc_normal_)
from .registry import register_model

__all__ = [
    "PretrainVisionTransformerEncoder",
    "PretrainVisionTransformerDecoder",
    "PretrainVisionTransformer",
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
        # one host sync, as x[~mask] has in the reference; every clip must keep the same number of tokens — the
        # reference fails loudly otherwise (x[~mask].reshape(B, -1, C) raises, mp:98), and so does this
        counts = (~mask).sum(1)
        lo, hi = int(counts.min()), int(counts.max())
        if lo != hi:
            raise ValueError(f"mask keeps between {lo} and {hi} tokens per clip; every clip must keep the same number")
        n_visible = lo
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


def masked_token_indices(mask, n_masked=None):
    """Row-major ids of the tokens x[mask] keeps (expand_pos_embed[mask], mp:286), int32 [B, n_masked]."""
    mask = mask.bool()
    if n_masked is None:
        n_masked = int(mask[0].sum())
    order = torch.argsort((~mask).to(torch.uint8), dim=1, stable=True)
    return order[:, :n_masked].to(torch.int32).contiguous(), n_masked


def _fold_norm_linear(norm, weight, bias, device):
    """LayerNorm folded into the Linear that follows it (see Block.packed): W' = W diag(gamma) in bf16,
    b' = b + W beta, colsum of the bf16 W'."""
    w = weight.detach().to(device=device, dtype=torch.float32)
    wf = w * norm.weight.detach().to(device=device, dtype=torch.float32)[None, :]
    bf = w @ norm.bias.detach().to(device=device, dtype=torch.float32)
    if bias is not None:
        bf = bf + bias.detach().to(device=device, dtype=torch.float32)
    wb = wf.to(torch.bfloat16).contiguous()
    return wb, bf.contiguous(), wb.float().sum(1).contiguous()


class PretrainVisionTransformerDecoder(nn.Module):
    """MAE decoder: blocks over all tokens, `norm`, pixel head on the last `return_token_num` tokens
    (modeling_pretrain.py:115-180)."""

    def __init__(self, patch_size=16, num_classes=768, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.,
                 qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, init_values=None, num_patches=196, tubelet_size=2, use_checkpoint=False,
                 use_flash_attn=True):
        super().__init__()
        self.num_classes = num_classes
        assert num_classes == 3 * tubelet_size * patch_size ** 2
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.use_checkpoint = use_checkpoint
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

    _init_weights = PretrainVisionTransformerEncoder._init_weights

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

    @torch.no_grad()
    def forward(self, x, return_token_num):
        """x [B, N, D] -> head(norm(x[:, -return_token_num:])) fp32 (all tokens when return_token_num <= 0), mp:164-178.
        Module-level path (stand-alone use); PretrainVisionTransformer.forward runs the fused stad_mae_forward."""
        _inference_only(self)
        if not isinstance(self.head, nn.Linear):
            raise NotImplementedError("decoder without a pixel head is never built by the reference (mp:234-251)")
        B, N, D = x.shape
        for blk in self.blocks:
            x = blk(x)
        h = _as_bf16_2d(x)
        w, b, cs = _fold_norm_linear(self.norm, self.head.weight, self.head.bias, h.device)
        y = _lib.ln_gemm(h, _lib.row_stats(h, self.norm.eps), w, b, cs).view(B, N, -1)
        return _lib.tail_rows_f32(y, return_token_num if return_token_num > 0 else N)


class _PreparedMae:
    """Packed weights of the whole pre-training model + the ctypes stad_mae_model pointing at them."""

    def __init__(self, owner, device):
        enc, dec = owner.encoder, owner.decoder
        self.enc = _PreparedModel(enc, device, enc.norm, None)
        keep = []
        blocks = (_lib.StadBlock * len(dec.blocks))()
        for i, blk in enumerate(dec.blocks):
            pk = blk.packed(device)
            keep.append(pk)
            for name, t in pk.items():
                setattr(blocks[i], name, t.data_ptr())
        Dd = dec.embed_dim
        w_e2d, b_e2d, cs_e2d = _fold_norm_linear(enc.norm, owner.encoder_to_decoder.weight,
                                                 owner.encoder_to_decoder.bias, device)
        w_pix, b_pix, cs_pix = _fold_norm_linear(dec.norm, dec.head.weight, dec.head.bias, device)
        pos_dec = owner.pos_embed.detach().to(device=device, dtype=torch.float32).reshape(-1, Dd).contiguous()
        mask_token = owner.mask_token.detach().to(device=device, dtype=torch.float32).reshape(Dd).contiguous()
        m = _lib.StadMaeModel()
        m.encoder = self.enc.model
        m.dec_dims = enc.patch_embed.stad_dims(depth=len(dec.blocks), heads=dec.num_heads,
                                               hidden=dec.blocks[0].mlp.fc1.out_features, num_classes=dec.num_classes)
        m.dec_dims.dim = Dd
        m.w_e2d, m.b_e2d, m.cs_e2d = w_e2d.data_ptr(), b_e2d.data_ptr(), cs_e2d.data_ptr()
        m.pos_dec, m.mask_token = pos_dec.data_ptr(), mask_token.data_ptr()
        m.dec_blocks = C.cast(blocks, C.POINTER(_lib.StadBlock))
        m.w_pix, m.b_pix, m.cs_pix = w_pix.data_ptr(), b_pix.data_ptr(), cs_pix.data_ptr()
        self.model = m
        self.device = device
        self.n_tokens = enc.patch_embed.num_patches
        self.num_classes = dec.num_classes
        self._keep = (keep, blocks, w_e2d, b_e2d, cs_e2d, w_pix, b_pix, cs_pix, pos_dec, mask_token)
        self._workspace = None
        self.last_launches = 0

    def run(self, inp, vis_idx, mask_idx, B, n_vis):
        lib = _lib.load()
        need = lib.stad_mae_workspace_bytes(C.byref(self.model), B, n_vis)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, self.n_tokens - n_vis, self.num_classes, dtype=torch.float32, device=self.device)
        rc = lib.stad_mae_forward(C.byref(self.model), C.byref(inp), _lib.ptr(vis_idx), _lib.ptr(mask_idx), B, n_vis,
                                  _lib.ptr(out), _lib.ptr(self._workspace), need, _lib.stream_ptr())
        self.last_launches = _lib.check(rc, "stad_mae_forward")
        return out


class PretrainVisionTransformer(nn.Module):
    """VideoMAE pre-training model: encoder over the visible tokens + light decoder predicting the pixels of the masked
    tubelets (modeling_pretrain.py:183-291), sm_100a forward."""

    def __init__(self,
                 img_size=224,
                 patch_size=16,
                 encoder_in_chans=3,
                 encoder_num_classes=0,
                 encoder_embed_dim=768,
                 encoder_depth=12,
                 encoder_num_heads=12,
                 decoder_num_classes=1536,
                 decoder_embed_dim=512,
                 decoder_depth=8,
                 decoder_num_heads=8,
                 mlp_ratio=4.,
                 qkv_bias=False,
                 qk_scale=None,
                 drop_rate=0.,
                 attn_drop_rate=0.,
                 drop_path_rate=0.,
                 norm_layer=nn.LayerNorm,
                 init_values=0.,
                 use_learnable_pos_emb=False,
                 use_flash_attn=True,
                 use_checkpoint=False,
                 tubelet_size=2,
                 num_classes=0,  # avoid the error from create_fn in timm
                 in_chans=0,  # avoid the error from create_fn in timm
                 ):
        super().__init__()
        self.encoder = PretrainVisionTransformerEncoder(
            img_size=img_size, patch_size=patch_size, in_chans=encoder_in_chans, num_classes=encoder_num_classes,
            embed_dim=encoder_embed_dim, depth=encoder_depth, num_heads=encoder_num_heads, mlp_ratio=mlp_ratio,
            qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
            drop_path_rate=drop_path_rate, norm_layer=norm_layer, init_values=init_values, tubelet_size=tubelet_size,
            use_checkpoint=use_checkpoint, use_learnable_pos_emb=use_learnable_pos_emb, use_flash_attn=use_flash_attn)
        self.decoder = PretrainVisionTransformerDecoder(
            patch_size=patch_size, num_patches=self.encoder.patch_embed.num_patches, num_classes=decoder_num_classes,
            embed_dim=decoder_embed_dim, depth=decoder_depth, num_heads=decoder_num_heads, mlp_ratio=mlp_ratio,
            qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
            drop_path_rate=drop_path_rate, norm_layer=norm_layer, init_values=init_values, tubelet_size=tubelet_size,
            use_checkpoint=use_checkpoint, use_flash_attn=use_flash_attn)
        self.encoder_to_decoder = nn.Linear(encoder_embed_dim, decoder_embed_dim, bias=False)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.pos_embed = get_sinusoid_encoding_table(self.encoder.patch_embed.num_patches, decoder_embed_dim)
        trunc_normal_(self.mask_token, std=.02)

    _init_weights = PretrainVisionTransformerEncoder._init_weights

    def get_num_layers(self):
        # the reference returns len(self.blocks), an attribute this class never has (mp:268-269); report the encoder's
        return len(self.encoder.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'mask_token'}

    def prepare(self, device=None):
        """Pack the weights for the kernels now (otherwise done lazily by the first forward after any weight change)."""
        _inference_only(self)
        if not isinstance(self.encoder.head, nn.Identity):
            raise NotImplementedError("encoder_num_classes > 0 is never used by the reference (mp:203)")
        self.encoder.blocks[0].attn._check_head_dim()
        self.decoder.blocks[0].attn._check_head_dim()
        device = torch.device(device or next(self.parameters()).device)
        sig = (_weights_signature(self), str(device))
        prep = getattr(self, "_stad_prepared", None)
        if prep is None or self._stad_sig != sig:
            _lib.init(device)
            prep = _PreparedMae(self, device)
            object.__setattr__(self, "_stad_prepared", prep)
            object.__setattr__(self, "_stad_sig", sig)
        return prep

    @torch.no_grad()
    def forward(self, x, mask, n_visible=None):
        """x [B, C, T, H, W], mask [B, N] bool (True = masked) -> [B, N_mask, 3*tubelet*patch^2] fp32: the predicted
        pixels of the masked tokens, row-major token order per clip (mp:276-291)."""
        if not x.is_cuda:
            raise RuntimeError("simple-tad_b200 runs on a CUDA (sm_100a) device only; got a CPU tensor")
        B, Cc, T, H, W = x.shape
        pe = self.encoder.patch_embed
        assert H == pe.img_size[0] and W == pe.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]})."
        if mask.shape != (B, pe.num_patches):
            raise ValueError(f"mask must be [{B}, {pe.num_patches}], got {tuple(mask.shape)}")
        prep = self.prepare(x.device)
        mask = mask.to(x.device)
        vis_idx, n_vis = visible_token_indices(mask, n_visible)
        mask_idx, _ = masked_token_indices(mask, pe.num_patches - n_vis)
        xb = prep.enc.input_bf16(x)
        inp = _lib.make_input(xb, _lib.STAD_IN_CLIPS)
        return prep.run(inp, vis_idx, mask_idx, B, n_vis)


_PRETRAIN_SPECS = {
    # reference factories modeling_pretrain.py:293-387: name -> encoder (dim, depth, heads), decoder (dim, heads)
    "pretrain_videomae_small_patch16_224": (384, 12, 6, 192, 3),
    "pretrain_videomae_base_patch16_224": (768, 12, 12, 384, 6),
    "pretrain_videomae_large_patch16_224": (1024, 24, 16, 512, 8),
    "pretrain_videomae_huge_patch16_224": (1280, 32, 16, 640, 8),
}


def _make_pretrain_factory(name, e_dim, e_depth, e_heads, d_dim, d_heads):
    def factory(pretrained=False, **kwargs):
        model = PretrainVisionTransformer(
            img_size=224, patch_size=16, encoder_embed_dim=e_dim, encoder_depth=e_depth, encoder_num_heads=e_heads,
            encoder_num_classes=0, decoder_num_classes=1536, decoder_embed_dim=d_dim, decoder_num_heads=d_heads,
            mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
        model.default_cfg = _cfg()
        if pretrained:  # mp:309-313
            checkpoint = torch.load(kwargs["init_ckpt"], map_location="cpu")
            model.load_state_dict(checkpoint["model"])
        return model
    factory.__name__ = factory.__qualname__ = name
    factory.__doc__ = (f"{name}: encoder D={e_dim} depth={e_depth} heads={e_heads}; decoder D={d_dim} heads={d_heads} "
                       "(reference factory of the same name; decoder_depth is a kwarg, default 8, the DAPT scripts pass 4).")
    return register_model(factory)


for _name, _spec in _PRETRAIN_SPECS.items():
    globals()[_name] = _make_pretrain_factory(_name, *_spec)
del _name, _spec
