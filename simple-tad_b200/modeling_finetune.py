"""B200-native drop-in for the reference's modeling_finetune.py (the Video-ViT classifier).

Same classes, constructor signatures, parameter names / shapes and factories as the reference
(modeling_finetune.py:37-398), so `load_state_dict` of a reference checkpoint works unchanged.  The math of
`forward` does not run in PyTorch: every module calls the hand-written sm_100a kernels of libstad.so through the C
ABI (include/stad.h).  Inference only (eval mode, no autograd); there is no PyTorch / CPU fallback.

    PatchEmbed         -> stad_patch_embed          (Conv3d as an im2col-free tcgen05 GEMM)
    Attention          -> stad_gemm_bias_residual + stad_attention (fused flash-style tcgen05 kernel)
    Mlp                -> stad_ln_gemm (bias+GELU epilogue) + stad_gemm_bias_residual
    Block              -> LayerNorm folded into the following GEMM, residual adds in the GEMM epilogues
    VisionTransformer  -> stad_vit_forward          (whole forward sequenced in C++, one ctypes call)
"""
import ctypes as C
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .registry import register_model

__all__ = [
    "Mlp", "Attention", "Block", "PatchEmbed", "DropPath", "get_sinusoid_encoding_table", "VisionTransformer", "_cfg",
    "vit_small_patch16_224", "vit_base_patch16_224", "vit_base_patch16_384", "vit_large_patch16_224",
    "vit_large_patch16_384", "vit_large_patch16_512", "vit_huge_patch16_224",
]


def _cfg(url='', **kwargs):
    """Same default_cfg dict as the reference (modeling_finetune.py:13-20)."""
    return {
        'url': url,
        'num_classes': 400, 'input_size': (3, 224, 224), 'pool_size': None,
        'crop_pct': .9, 'interpolation': 'bicubic',
        'mean': (0.5, 0.5, 0.5), 'std': (0.5, 0.5, 0.5),
        **kwargs
    }


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _inference_only(module):
    if module.training:
        raise NotImplementedError(
            f"{type(module).__name__}: simple-tad_b200 implements the inference forward only; call .eval() "
            "(training / backward is outside the accelerated path)")


def _as_bf16_2d(x):
    """[B, N, C] (fp32 / fp16 / bf16, CUDA) -> contiguous bf16 [B*N, C]."""
    if not x.is_cuda:
        raise RuntimeError("simple-tad_b200 modules run on a CUDA (sm_100a) device only; got a CPU tensor")
    B, N, Cc = x.shape
    return x.reshape(B * N, Cc).to(torch.bfloat16).contiguous()


class DropPath(nn.Module):
    """Stochastic depth (modeling_finetune.py:23-34). Identity at inference, which is all this package runs."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        _inference_only(self)
        return x

    def extra_repr(self) -> str:
        return 'p={}'.format(self.drop_prob)


class Mlp(nn.Module):
    """fc2(GELU_erf(fc1(x))) (modeling_finetune.py:37-54). GELU and the fc1 bias run in the fc1 GEMM epilogue."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("only nn.GELU (exact erf) is implemented; the reference never uses another act")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    @torch.no_grad()
    def _pack(self, device):
        f32 = dict(dtype=torch.float32, device=device)
        return dict(w1=self.fc1.weight.detach().to(device=device, dtype=torch.bfloat16).contiguous(),
                    b1=self.fc1.bias.detach().to(**f32).contiguous(),
                    zeros=torch.zeros(self.fc1.out_features, **f32),
                    w2=self.fc2.weight.detach().to(device=device, dtype=torch.bfloat16).contiguous(),
                    b2=self.fc2.bias.detach().to(**f32).contiguous(), ident={})

    def forward(self, x):
        _inference_only(self)
        B, N, _ = x.shape
        a = _as_bf16_2d(x)
        M = a.shape[0]
        pk = _cached_pack(self, a.device, self._pack)
        # identity LayerNorm statistics (mean 0, rstd 1) turn the LN-fold epilogue into plain bias + GELU
        ident = pk["ident"].get(M)
        if ident is None:
            ident = torch.zeros(M, 2, dtype=torch.float32, device=a.device)
            ident[:, 1] = 1.0
            pk["ident"].clear()
            pk["ident"][M] = ident
        h = _lib.ln_gemm(a, ident, pk["w1"], pk["b1"], pk["zeros"], gelu=True)
        y = _lib.gemm_bias_residual(h, pk["w2"], pk["b2"])
        return y.view(B, N, -1).to(x.dtype)


class Attention(nn.Module):
    """Joint space-time multi-head attention (modeling_finetune.py:57-134).

    `use_flash_attn` is accepted for signature parity; both values run the same fused sm_100a kernel, whose results
    match `_naive_attn` (mf:86-106) within bf16 tolerance and which, like `_flash_attn` (mf:108-130), never
    materialises the N x N score matrix."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.,
                 attn_head_dim=None, use_flash_attn=False, causal=False):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if attn_head_dim is not None:
            head_dim = attn_head_dim
        self.head_dim = head_dim
        if causal:
            raise NotImplementedError("causal attention is never used on this path (mf:146)")
        all_head_dim = head_dim * self.num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, all_head_dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(all_head_dim))
            self.v_bias = nn.Parameter(torch.zeros(all_head_dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(all_head_dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.use_flash_attn = use_flash_attn

    def _check_head_dim(self):
        if self.head_dim != 64:
            raise NotImplementedError(f"head_dim={self.head_dim}: the sm_100a attention kernel is specialised for the "
                                      "head dim 64 of VideoMAE ViT-S/B/L (ViT-H uses 80 and is not on this path)")

    def packed_qkv_bias(self):
        """cat(q_bias, 0, v_bias) — K has no bias (mf:88-90); built once per weight prep, not per forward."""
        if self.q_bias is None:
            return None
        return torch.cat((self.q_bias.detach(), torch.zeros_like(self.v_bias), self.v_bias.detach())).float()

    @torch.no_grad()
    def _pack(self, device):
        bias = self.packed_qkv_bias()
        return dict(w_qkv=self.qkv.weight.detach().to(device=device, dtype=torch.bfloat16).contiguous(),
                    b_qkv=None if bias is None else bias.to(device).contiguous(),
                    w_proj=self.proj.weight.detach().to(device=device, dtype=torch.bfloat16).contiguous(),
                    b_proj=self.proj.bias.detach().to(device=device, dtype=torch.float32).contiguous())

    def forward(self, x):
        _inference_only(self)
        self._check_head_dim()
        B, N, _ = x.shape
        a = _as_bf16_2d(x)
        pk = _cached_pack(self, a.device, self._pack)
        qkv = _lib.gemm_bias_residual(a, pk["w_qkv"], pk["b_qkv"])
        ctx = _lib.attention(qkv.view(B, N, 3, self.num_heads, 64), scale=self.scale)
        y = _lib.gemm_bias_residual(ctx.view(B * N, -1), pk["w_proj"], pk["b_proj"])
        return y.view(B, N, -1).to(x.dtype)


class Block(nn.Module):
    """x + Attn(LN1(x)); x + Mlp(LN2(x))  (modeling_finetune.py:137-166)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 attn_head_dim=None, use_flash_attn=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(
            dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
            attn_drop=attn_drop, proj_drop=drop, attn_head_dim=attn_head_dim, use_flash_attn=use_flash_attn,
            causal=False)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        # the reference evaluates `init_values > 0` and therefore needs a number (mf:153); None is treated as 0 here
        if init_values is not None and init_values > 0:
            self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
            self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        else:
            self.gamma_1, self.gamma_2 = None, None
        for n in (self.norm1, self.norm2):
            if not isinstance(n, nn.LayerNorm):
                raise NotImplementedError("norm_layer must be nn.LayerNorm (it is folded into the following GEMM)")

    @torch.no_grad()
    def packed(self, device):
        """Weights of this block in the layout the kernels consume (see stad_block in include/stad.h):
        LayerNorm gamma folded into the next weight, beta into its bias, per-row sums of the bf16 weight for the
        mean correction, layer-scale gamma (if any) folded into proj / fc2."""
        f32 = dict(dtype=torch.float32, device=device)

        def fold(norm, w, b):
            w = w.detach().to(**f32)
            wf = w * norm.weight.detach().to(**f32)[None, :]
            bf = w @ norm.bias.detach().to(**f32)
            if b is not None:
                bf = bf + b.to(**f32)
            wb = wf.to(torch.bfloat16).contiguous()
            return wb, bf.contiguous(), wb.float().sum(1).contiguous()

        def scaled(lin, gamma):
            w = lin.weight.detach().to(**f32)
            b = lin.bias.detach().to(**f32)
            if gamma is not None:
                g = gamma.detach().to(**f32)
                w, b = w * g[:, None], b * g
            return w.to(torch.bfloat16).contiguous(), b.contiguous()

        w_qkv, b_qkv, cs_qkv = fold(self.norm1, self.attn.qkv.weight, self.attn.packed_qkv_bias())
        w_fc1, b_fc1, cs_fc1 = fold(self.norm2, self.mlp.fc1.weight, self.mlp.fc1.bias.detach())
        w_proj, b_proj = scaled(self.attn.proj, self.gamma_1)
        w_fc2, b_fc2 = scaled(self.mlp.fc2, self.gamma_2)
        return dict(w_qkv=w_qkv, b_qkv=b_qkv, cs_qkv=cs_qkv, w_proj=w_proj, b_proj=b_proj, w_fc1=w_fc1, b_fc1=b_fc1,
                    cs_fc1=cs_fc1, w_fc2=w_fc2, b_fc2=b_fc2)

    def forward(self, x):
        _inference_only(self)
        self.attn._check_head_dim()
        B, N, D = x.shape
        h = _as_bf16_2d(x)
        pk = _cached_pack(self, h.device, self.packed)
        eps1, eps2 = self.norm1.eps, self.norm2.eps
        st = _lib.row_stats(h, eps1)
        qkv = _lib.ln_gemm(h, st, pk["w_qkv"], pk["b_qkv"], pk["cs_qkv"])
        ctx = _lib.attention(qkv.view(B, N, 3, self.attn.num_heads, 64), scale=self.attn.scale)
        h, st = _lib.gemm_bias_residual_stats(ctx.view(B * N, D), pk["w_proj"], pk["b_proj"], h, eps2)
        hid = _lib.ln_gemm(h, st, pk["w_fc1"], pk["b_fc1"], pk["cs_fc1"], gelu=True)
        h = _lib.gemm_bias_residual(hid, pk["w_fc2"], pk["b_fc2"], residual=h)
        return h.view(B, N, D).to(x.dtype)


class PatchEmbed(nn.Module):
    """Video to tubelet-patch embedding (modeling_finetune.py:169-191). `proj` keeps the Conv3d parameter layout
    ([D, C, tubelet, p, p] weight + bias) for checkpoint compatibility; the forward is an im2col-free GEMM."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.tubelet_size = int(tubelet_size)
        num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * (num_frames // self.tubelet_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = num_patches
        self.num_frames = num_frames
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv3d(in_channels=in_chans, out_channels=embed_dim,
                              kernel_size=(self.tubelet_size, patch_size[0], patch_size[1]),
                              stride=(self.tubelet_size, patch_size[0], patch_size[1]))

    def stad_dims(self, depth=0, heads=0, hidden=0, num_classes=0):
        return _lib.make_dims(img_h=self.img_size[0], img_w=self.img_size[1], patch=self.patch_size[0],
                              tubelet=self.tubelet_size, frames=self.num_frames, in_chans=self.in_chans,
                              dim=self.embed_dim, depth=depth, heads=heads, hidden=hidden, num_classes=num_classes)

    def forward(self, x, **kwargs):
        _inference_only(self)
        B, Cc, T, H, W = x.shape
        # FIXME of the reference kept: size constraints are not relaxed (mf:187-189)
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        if T != self.num_frames:
            raise ValueError(f"PatchEmbed: expected {self.num_frames} frames, got {T}")
        if not x.is_cuda:
            raise RuntimeError("simple-tad_b200 modules run on a CUDA (sm_100a) device only; got a CPU tensor")
        xb = x.contiguous()
        xb = _lib.cast_f32_bf16(xb) if xb.dtype == torch.float32 else xb.to(torch.bfloat16)
        D = self.embed_dim

        def pack(device):
            w = self.proj.weight.detach().reshape(D, -1).to(device=device, dtype=torch.bfloat16).contiguous()
            bias = (self.proj.bias.detach().to(device=device, dtype=torch.float32) if self.proj.bias is not None
                    else torch.zeros(D, device=device))
            return w, bias[None, :].expand(self.num_patches, D).contiguous()  # no position table at module level
        w, pos_bias = _cached_pack(self, x.device, pack)
        out = _lib.patch_embed(xb, w, pos_bias, self.stad_dims(), B, self.num_patches)
        return out.view(B, self.num_patches, D).to(x.dtype)


def get_sinusoid_encoding_table(n_position, d_hid):
    """Sinusoid position encoding table, identical arithmetic to the reference (modeling_finetune.py:195-205):
    float64 angles over the flat token index, sin on even / cos on odd channels, cast to fp32, shape [1, N, D].
    (Vectorised; the reference builds the same numbers with a Python list comprehension.)"""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    hid = np.arange(d_hid)[None, :]
    sinusoid_table = pos / np.power(10000, 2 * (hid // 2) / d_hid)
    sinusoid_table[:, 0::2] = np.sin(sinusoid_table[:, 0::2])
    sinusoid_table[:, 1::2] = np.cos(sinusoid_table[:, 1::2])
    return torch.tensor(sinusoid_table, dtype=torch.float, requires_grad=False).unsqueeze(0)


class _PreparedModel:
    """Device-side packed weights + the ctypes stad_model that points at them (built once per weight version)."""

    def __init__(self, owner, device, final_norm, head, reduction=_lib.STAD_REDUCE_MEAN, cls_token=None):
        pe = owner.patch_embed
        D = owner.embed_dim
        keep = []
        blocks = (_lib.StadBlock * len(owner.blocks))()
        for i, blk in enumerate(owner.blocks):
            pk = blk.packed(device)
            keep.append(pk)
            for name, t in pk.items():
                setattr(blocks[i], name, t.data_ptr())
        w_patch = pe.proj.weight.detach().to(device).reshape(D, -1).to(torch.bfloat16).contiguous()
        pos = owner.pos_embed.detach().to(device=device, dtype=torch.float32).reshape(-1, D)
        if pos.shape[0] != pe.num_patches:
            raise ValueError(f"pos_embed has {pos.shape[0]} rows for {pe.num_patches} patches")
        bias = pe.proj.bias.detach().to(device=device, dtype=torch.float32) if pe.proj.bias is not None else 0.0
        pos_bias = (pos + bias).contiguous()
        norm_g = final_norm.weight.detach().to(device=device, dtype=torch.float32).contiguous()
        norm_b = final_norm.bias.detach().to(device=device, dtype=torch.float32).contiguous()
        num_classes = 0
        w_head = b_head = None
        if head is not None:
            num_classes = head.out_features
            w_head = head.weight.detach().to(device=device, dtype=torch.float32).contiguous()
            b_head = head.bias.detach().to(device=device, dtype=torch.float32).contiguous()
        heads = owner.num_heads
        hidden = owner.blocks[0].mlp.fc1.out_features
        self.dims = pe.stad_dims(depth=len(owner.blocks), heads=heads, hidden=hidden, num_classes=num_classes)
        m = _lib.StadModel()
        m.dims = self.dims
        m.w_patch = w_patch.data_ptr()
        m.pos_bias = pos_bias.data_ptr()
        m.blocks = C.cast(blocks, C.POINTER(_lib.StadBlock))
        m.norm_g = norm_g.data_ptr()
        m.norm_b = norm_b.data_ptr()
        m.w_head = w_head.data_ptr() if w_head is not None else None
        m.b_head = b_head.data_ptr() if b_head is not None else None
        m.eps = float(final_norm.eps)
        m.attn_scale = float(owner.blocks[0].attn.scale)
        m.reduction = int(reduction)
        if cls_token is not None:
            cls_token = cls_token.detach().to(device=device, dtype=torch.float32).reshape(D).contiguous()
            m.cls_token = cls_token.data_ptr()
        self.reduction = int(reduction)
        self.cls_rows = 0 if cls_token is None else 1
        self.model = m
        self.device = device
        self.num_classes = num_classes
        self.embed_dim = D
        self.n_tokens = pe.num_patches
        self._keep = (keep, blocks, w_patch, pos_bias, norm_g, norm_b, w_head, b_head, cls_token)
        self._workspace = None
        self._ws_key = None
        self._in_bf16 = None

    def graph_token(self):
        """What a captured CUDA graph of a forward has baked in: this prepared model (the packed weights) and the raw
        pointers of its workspace and input-cast buffer.  A graph is only replayable while the token is unchanged
        (another caller with a larger batch reallocates the workspace; a weight change builds a new prepared model)."""
        return (id(self), self._workspace.data_ptr() if self._workspace is not None else 0,
                self._in_bf16.data_ptr() if self._in_bf16 is not None else 0)

    def workspace(self, B, n_tok):
        need = _lib.load().stad_workspace_bytes(C.byref(self.dims), B, n_tok)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace, need

    def input_bf16(self, x):
        """fp32 clips -> bf16 through stad_cast_f32_bf16 into a reused buffer (autocast's input cast, eff:428)."""
        if x.dtype == torch.bfloat16:
            return x.contiguous()
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        if self._in_bf16 is None or self._in_bf16.numel() < x.numel():
            self._in_bf16 = torch.empty(x.numel(), dtype=torch.bfloat16, device=self.device)
        buf = self._in_bf16[: x.numel()].view(x.shape)
        _lib.check(_lib.load().stad_cast_f32_bf16(_lib.ptr(x), _lib.ptr(buf), x.numel(), _lib.stream_ptr()),
                   "stad_cast_f32_bf16")
        return buf

    def run(self, inp, B, n_tok, tok_idx=None, want=("logits",)):
        """One stad_vit_forward call. `want`: any of logits / probs / features (classifier) or tokens (encoder).
        Returns a dict of fp32 tensors."""
        lib = _lib.load()
        rows = n_tok + self.cls_rows                       # rows per clip in the residual stream
        ws, need = self.workspace(B, rows)
        per_tok = (rows,) if self.reduction == _lib.STAD_REDUCE_NONE else ()  # 'none': one result per token (mf:330)
        shapes = {"logits": (B,) + per_tok + (self.num_classes,), "probs": (B,) + per_tok + (self.num_classes,),
                  "features": (B,) + per_tok + (self.embed_dim,), "tokens": (B, rows, self.embed_dim)}
        res = {k: torch.empty(shapes[k], dtype=torch.float32, device=self.device) for k in want}
        outs = _lib.StadOutputs(*[res[k].data_ptr() if k in res else None for k in ("logits", "probs", "features", "tokens")])
        rc = lib.stad_vit_forward(C.byref(self.model), C.byref(inp), _lib.ptr(tok_idx), B, n_tok, C.byref(outs),
                                  _lib.ptr(ws), need, _lib.stream_ptr())
        self.last_launches = _lib.check(rc, "stad_vit_forward")
        return res


# Walking module.parameters() costs ~0.4 ms for a ViT-B (162 tensors behind a de-duplicating generator) — half a
# batch-1 forward.  The tensor LIST is therefore cached per module and rebuilt only when some module somewhere registered
# a parameter, buffer or sub-module since (global registration hooks bump an epoch); the signature itself — (data_ptr,
# version) of every tensor, which moves on load_state_dict / in-place updates / .to() — is re-read every time (~40 us).
_STRUCT_EPOCH = [0]


def _bump_struct_epoch(*args, **kwargs):
    _STRUCT_EPOCH[0] += 1


torch.nn.modules.module.register_module_parameter_registration_hook(_bump_struct_epoch)
torch.nn.modules.module.register_module_buffer_registration_hook(_bump_struct_epoch)
torch.nn.modules.module.register_module_module_registration_hook(_bump_struct_epoch)


def _weights_signature(module):
    cache = module.__dict__.get("_stad_tensor_list")
    if cache is None or cache[0] != _STRUCT_EPOCH[0]:
        cache = (_STRUCT_EPOCH[0], list(module.parameters()) + list(module.buffers()))
        object.__setattr__(module, "_stad_tensor_list", cache)
    return tuple((t.data_ptr(), t._version) for t in cache[1])


def _cached_pack(module, device, build):
    """build(device) -> packed weights of `module`, cached on the module until one of its tensors changes (stand-alone
    Block / Mlp / Attention / PatchEmbed calls — modeling_pretrain composes Blocks this way, mp:8 — no longer re-fold and
    re-cast their weights on every forward)."""
    sig = (_weights_signature(module), str(device))
    hit = module.__dict__.get("_stad_pack")
    if hit is None or hit[0] != sig:
        hit = (sig, build(device))
        object.__setattr__(module, "_stad_pack", hit)
    return hit[1]


class _StadBackbone(nn.Module):
    """Shared plumbing of the classifier and the pre-training encoder: lazy weight preparation + cache."""

    def _prepared_for(self, device, final_norm, head, reduction=_lib.STAD_REDUCE_MEAN, cls_token=None):
        sig = (_weights_signature(self), str(device), int(reduction))
        prep = getattr(self, "_stad_prepared", None)
        if prep is None or self._stad_sig != sig:
            _lib.init(device)
            prep = _PreparedModel(self, device, final_norm, head, reduction, cls_token)
            object.__setattr__(self, "_stad_prepared", prep)
            object.__setattr__(self, "_stad_sig", sig)
        return prep

    def prepare(self):
        """Pack the weights for the kernels now (otherwise done lazily by the first forward after any weight change)."""
        raise NotImplementedError


class VisionTransformer(_StadBackbone):
    """Vision Transformer for video clips (modeling_finetune.py:208-335), sm_100a forward."""

    def __init__(self,
                 img_size=224,
                 patch_size=16,
                 in_chans=3,
                 num_classes=1000,
                 embed_dim=768,
                 depth=12,
                 num_heads=12,
                 mlp_ratio=4.,
                 qkv_bias=False,
                 qk_scale=None,
                 fc_drop_rate=0.,
                 drop_rate=0.,
                 attn_drop_rate=0.,
                 drop_path_rate=0.,
                 norm_layer=nn.LayerNorm,
                 init_values=0.,
                 use_learnable_pos_emb=False,
                 use_flash_attn=True,
                 init_scale=0.,
                 all_frames=16,
                 tubelet_size=2,
                 use_checkpoint=False,
                 final_reduction="fc_norm"):
        super().__init__()
        self.num_classes = num_classes
        self.num_heads = num_heads
        self.num_features = self.embed_dim = embed_dim  # num_features for consistency with other models
        self.tubelet_size = tubelet_size
        self.num_frames = all_frames
        self.patch_embed = PatchEmbed(
            img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, num_frames=all_frames,
            tubelet_size=self.tubelet_size)
        num_patches = self.patch_embed.num_patches
        self.use_checkpoint = use_checkpoint  # activation checkpointing is a training feature: ignored at inference

        self.pos_embed = self._build_pos_embed(num_patches, embed_dim, use_learnable_pos_emb)

        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]  # stochastic depth decay rule
        self.blocks = nn.ModuleList([
            Block(
                dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                init_values=init_values, use_flash_attn=use_flash_attn)
            for i in range(depth)])
        assert final_reduction in ("fc_norm", "cls", 'none', None)
        self.final_reduction = final_reduction
        self.norm = nn.Identity() if final_reduction == "fc_norm" else norm_layer(embed_dim)
        self.fc_norm = norm_layer(embed_dim) if final_reduction == "fc_norm" else None
        self.fc_dropout = nn.Dropout(p=fc_drop_rate) if fc_drop_rate > 0 else nn.Identity()
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

        if use_learnable_pos_emb:
            trunc_normal_(self.pos_embed, std=.02)
        self._init_extra_tokens()

        if hasattr(self.head, "weight"):
            trunc_normal_(self.head.weight, std=.02)
        self.apply(self._init_weights)

        if hasattr(self.head, "weight"):
            self.head.weight.data.mul_(init_scale)
            self.head.bias.data.mul_(init_scale)

    def _build_pos_embed(self, num_patches, embed_dim, learnable):
        """The position table of the model (mf:249-253).  The sibling architectures (other_models/) override this."""
        if learnable:
            return nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        # plain tensor attribute, absent from the state_dict, exactly as in the reference (mf:253)
        return get_sinusoid_encoding_table(num_patches, embed_dim)

    def _init_extra_tokens(self):
        """Hook for parameters a sibling adds between the blocks and the head (MVD's class token)."""

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
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

    # ------------------------------------------------------------------------------------------ sm_100a forward
    def _check_supported(self):
        _inference_only(self)
        if not isinstance(self.head, nn.Linear):
            raise NotImplementedError("num_classes=0 on VisionTransformer: use PretrainVisionTransformerEncoder for features")
        self.blocks[0].attn._check_head_dim()

    def _reduction(self):
        """final_reduction -> (STAD_REDUCE_*, the LayerNorm the kernels apply).  mf:323-330: 'fc_norm' pools and then
        applies fc_norm (norm is Identity); 'cls' / anything else apply norm and keep token 0 / every token."""
        if self.final_reduction == "fc_norm":
            return _lib.STAD_REDUCE_MEAN, self.fc_norm
        if self.final_reduction == "cls":
            return _lib.STAD_REDUCE_CLS, self.norm
        return _lib.STAD_REDUCE_NONE, self.norm

    def _cls_token(self):
        return None  # the simple-tad ViT has no class token; the MVD sibling overrides this

    def prepare(self, device=None):
        self._check_supported()
        device = device or next(self.parameters()).device
        reduction, norm = self._reduction()
        return self._prepared_for(torch.device(device), norm, self.head, reduction, self._cls_token())

    def _run(self, x, want):
        self._check_supported()
        if x.dim() != 5:
            raise ValueError(f"expected clips [B, C, T, H, W], got {tuple(x.shape)}")
        B, Cc, T, H, W = x.shape
        pe = self.patch_embed
        assert H == pe.img_size[0] and W == pe.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]})."
        if T != self.num_frames or Cc != pe.in_chans:
            raise ValueError(f"expected [B, {pe.in_chans}, {self.num_frames}, H, W] clips, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("simple-tad_b200 runs on a CUDA (sm_100a) device only; got a CPU tensor")
        prep = self.prepare(x.device)
        xb = prep.input_bf16(x)
        inp = _lib.make_input(xb, _lib.STAD_IN_CLIPS)
        return prep.run(inp, B, pe.num_patches, want=want)

    @torch.no_grad()
    def forward_windows(self, frames, start=0, count=None, stride=1, frame_step=1, starts=None, reuse_tubelets=True):
        """Sliding-window inference straight from a resident frame buffer (ri:69-109, dota.py:204-223):
        frames [F, C, H, W] (fp32 or bf16, already normalised); frame t of window b is frames[start + b*stride +
        t*frame_step] (frame_step = orig_fps // target_fps, dataset/sequencing.py:45-58; 1 = consecutive frames).
        starts: optional int32 CUDA tensor [count] with the first frame of every window instead of start + b*stride —
        one batch can then hold windows of several videos laid end to end in `frames` (final_test, eff:385-463).
        reuse_tubelets: embed every distinct tubelet of the batch once and assemble the overlapping windows from them
        (stad_input.tubelet_reuse; the embedding is rounded to bf16 before the position add, so results match the clips
        path within tolerance; False keeps them bit-identical to it).
        Returns (logits, probs), each [count, num_classes], without materialising the [count, C, T, H, W] clips."""
        self._check_supported()
        if frames.dim() != 4 or not frames.is_cuda:
            raise ValueError(f"expected CUDA frames [F, C, H, W], got {tuple(frames.shape)} on {frames.device}")
        pe = self.patch_embed
        # the tensor map is built from the MODEL's geometry: frames of another size would be read with the wrong strides
        assert frames.shape[2] == pe.img_size[0] and frames.shape[3] == pe.img_size[1], \
            f"Input image size ({frames.shape[2]}*{frames.shape[3]}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]})."
        if frames.shape[1] != pe.in_chans:
            raise ValueError(f"expected frames [F, {pe.in_chans}, H, W], got {tuple(frames.shape)}")
        F_ = frames.shape[0]
        T = self.num_frames
        if frame_step < 1 or stride < 1:
            raise ValueError(f"stride={stride} and frame_step={frame_step} must be >= 1")
        span = (T - 1) * frame_step + 1                      # frames a window covers (sequencing.py:48-50)
        if starts is not None:
            if starts.dim() != 1 or starts.dtype != torch.int32 or starts.device != frames.device:
                raise ValueError("starts must be a 1-D int32 tensor on the device of the frames")
            count = starts.numel() if count is None else count
            if count > starts.numel():
                raise ValueError(f"count={count} but only {starts.numel()} window starts were given")
        elif count is None:
            count = (F_ - span - start) // stride + 1
        if count < 1 or F_ < span:
            raise ValueError(f"{F_} frames hold no window of {T} frames (frame step {frame_step}) from start={start}")
        prep = self.prepare(frames.device)
        fb = prep.input_bf16(frames)
        inp = _lib.make_input(fb, _lib.STAD_IN_FRAMES, n_frames=F_, start=start, stride=stride, frame_step=frame_step,
                              window_starts=starts, tubelet_reuse=reuse_tubelets)
        res = prep.run(inp, count, pe.num_patches, want=("logits", "probs"))
        return res["logits"], res["probs"]

    @torch.no_grad()
    def forward_features(self, x):
        """fc_norm(mean over tokens) [B, D] (mf:308-326); 'cls': norm(x)[:, 0]; 'none': norm(x) [B, N, D] (mf:327-330)."""
        return self._run(x, want=("logits", "features"))["features"]

    @torch.no_grad()
    def forward(self, x):
        """logits [B, num_classes] (mf:332-335); [B, N, num_classes] with final_reduction='none'."""
        return self._run(x, want=("logits",))["logits"]

    @torch.no_grad()
    def forward_probs(self, x):
        """(logits, softmax(logits)) in one pass — VisionTransformerInfer.forward (ris:378-382) / ri:107."""
        res = self._run(x, want=("logits", "probs"))
        return res["logits"], res["probs"]


# Factories of the reference (modeling_finetune.py:338-398): name -> (img_size, embed_dim, depth, num_heads).
# All use patch 16, mlp_ratio 4, qkv_bias=True and LayerNorm(eps=1e-6).
_FACTORY_SPECS = {
    "vit_small_patch16_224": (224, 384, 12, 6),
    "vit_base_patch16_224": (224, 768, 12, 12),
    "vit_base_patch16_384": (384, 768, 12, 12),
    "vit_large_patch16_224": (224, 1024, 24, 16),
    "vit_large_patch16_384": (384, 1024, 24, 16),
    "vit_large_patch16_512": (512, 1024, 24, 16),
    "vit_huge_patch16_224": (224, 1280, 32, 16),
}


def _make_factory(name, img_size, embed_dim, depth, num_heads):
    def factory(pretrained=False, **kwargs):
        kwargs.setdefault("img_size", img_size)
        model = VisionTransformer(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4,
                                  qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
        model.default_cfg = _cfg()
        return model
    factory.__name__ = factory.__qualname__ = name
    factory.__doc__ = f"{name}: img {img_size}, D={embed_dim}, depth={depth}, heads={num_heads} (reference factory of the same name)."
    return register_model(factory)


for _name, _spec in _FACTORY_SPECS.items():
    globals()[_name] = _make_factory(_name, *_spec)
del _name, _spec
