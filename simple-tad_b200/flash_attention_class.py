"""Drop-in for the reference's flash_attention_class.FlashAttention (flash_attention_class.py:10-71), backed by the
hand-written sm_100a attention kernel (stad_attention) instead of flash_attn_varlen_qkvpacked_func."""
import torch
import torch.nn as nn

from . import _lib


class FlashAttention(nn.Module):
    """Scaled dot product attention with softmax over packed qkv.

    softmax_scale: temperature (default 1/sqrt(head_dim), computed at runtime, fac:13-15)
    attention_dropout: accepted for signature parity; must be 0 at inference (fac:48 passes 0.0 in eval)."""

    def __init__(self, softmax_scale=None, attention_dropout=0.0, device=None, dtype=None):
        super().__init__()
        self.softmax_scale = softmax_scale
        self.dropout_p = attention_dropout

    def forward(self, qkv, key_padding_mask=None, causal=False, cu_seqlens=None, max_s=None, need_weights=False):
        """qkv: (B, S, 3, H, D) -> (out (B, S, H, D), None)     [fac:26-51, the key_padding_mask=None branch]"""
        assert not need_weights                                   # fac:35
        assert qkv.is_cuda                                        # fac:37
        if self.training and self.dropout_p > 0:
            raise NotImplementedError("attention dropout is a training feature; this path is inference-only")
        if causal:
            raise NotImplementedError("causal attention is never used on the Video-ViT path (mf:146 passes causal=False)")
        if key_padding_mask is not None or cu_seqlens is not None:
            # dead on this path, and broken in the reference against flash-attn >= 2.6 (fac:55 unpacks 4 of 5 values)
            raise NotImplementedError("variable-length / padded attention is not part of the Video-ViT path")
        if qkv.dim() != 5 or qkv.shape[2] != 3:
            raise ValueError(f"qkv must be (B, S, 3, H, D), got {tuple(qkv.shape)}")
        B, S, _, H, D = qkv.shape
        in_dtype = qkv.dtype
        q = qkv.to(torch.bfloat16).contiguous()
        out = _lib.attention(q, scale=self.softmax_scale if self.softmax_scale is not None else D ** -0.5)
        out = out.view(B, S, H, D)
        return (out if in_dtype == torch.bfloat16 else out.to(in_dtype)), None
