"""CPU restatement (torch fp32, functional) of the reference Video-ViT forward.  TEST INFRASTRUCTURE ONLY — see
oracle/__init__.py.  Parity pinned by tests/golden/ (outputs of the unmodified reference, oracle/make_golden.py).

Citations are file:line in /root/reference (mf = modeling_finetune.py, mp = modeling_pretrain.py,
ris = run_inference_simple.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import synth

LN_EPS = 1e-6  # norm_layer=partial(nn.LayerNorm, eps=1e-6), mf:342


def sinusoid_table(n_position, d_hid):
    """mf:195-205: angle[p, j] = p / 10000^(2*(j//2)/d) in float64 over the FLAT token index p; sin on even j, cos on
    odd j; cast to fp32.  Shape [1, n_position, d_hid]."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000.0, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.tensor(table, dtype=torch.float).unsqueeze(0)


def patch_embed(sd, x):
    """mf:185-191: Conv3d(k = s = (2,16,16)) then flatten(2).transpose(1,2): tokens ordered (t', h', w')."""
    y = F.conv3d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=(synth.TUBELET, synth.PATCH, synth.PATCH))
    return y.flatten(2).transpose(1, 2)


def attention(sd, p, x, heads):
    """Attention._naive_attn, mf:86-106: K has no bias; q scaled by head_dim^-0.5; softmax(q k^T) v; proj."""
    B, N, C = x.shape
    bias = torch.cat((sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.v_bias"]), sd[p + "attn.v_bias"]))  # mf:90
    qkv = F.linear(x, sd[p + "attn.qkv.weight"], bias).reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)     # mf:92-93
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (q.shape[-1] ** -0.5)                                                                             # mf:67,96
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)                                                           # mf:97-100
    out = (attn @ v).transpose(1, 2).reshape(B, N, -1)                                                         # mf:103
    return F.linear(out, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])                                  # mf:104


def mlp(sd, p, x):
    """Mlp.forward, mf:47-54: fc2(GELU_erf(fc1(x)))."""
    return F.linear(F.gelu(F.linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])),
                    sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def block(sd, i, x, heads):
    """Block.forward with gamma_1 None (init_values=0), mf:159-162."""
    p = f"blocks.{i}."
    D = x.shape[-1]
    x = x + attention(sd, p, F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS), heads)
    x = x + mlp(sd, p, F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS))
    return x


def _depth(sd):
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


@torch.no_grad()
def vit_forward(sd, x, heads, return_hidden=False):
    """VisionTransformer.forward with final_reduction='fc_norm', mf:308-335.  x fp32 [B,3,16,224,224] -> logits.
    return_hidden: also the residual stream after patch-embed+pos and after every block."""
    D = sd["fc_norm.weight"].shape[0]
    h = patch_embed(sd, x)                                            # mf:309
    h = h + sinusoid_table(h.shape[1], D)                             # mf:312-313
    hidden = [h]
    for i in range(_depth(sd)):                                       # mf:320-321
        h = block(sd, i, h, heads)
        hidden.append(h)
    pooled = F.layer_norm(h.mean(1), (D,), sd["fc_norm.weight"], sd["fc_norm.bias"], LN_EPS)  # mf:323-326
    logits = F.linear(pooled, sd["head.weight"], sd["head.bias"])     # mf:334
    return (logits, hidden) if return_hidden else logits


@torch.no_grad()
def vit_probs(sd, x, heads):
    """VisionTransformerInfer.forward, ris:378-382: softmax over the two logits."""
    return vit_forward(sd, x, heads).softmax(-1)


@torch.no_grad()
def encoder_forward(sd, x, mask, heads):
    """PretrainVisionTransformerEncoder.forward_features, mp:91-108: embed ALL tokens, add pos, keep x[~mask] in
    row-major order, run the blocks, apply `norm` to every visible token (head = Identity, mp:60,112)."""
    D = sd["norm.weight"].shape[0]
    h = patch_embed(sd, x)
    h = h + sinusoid_table(h.shape[1], D)                             # mp:95
    B, _, C = h.shape
    h = h[~mask].reshape(B, -1, C)                                    # mp:98
    for i in range(_depth(sd)):
        h = block(sd, i, h, heads)
    return F.layer_norm(h, (D,), sd["norm.weight"], sd["norm.bias"], LN_EPS)  # mp:107


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@torch.no_grad()
def decoder_forward(sd, x, return_token_num, heads):
    """PretrainVisionTransformerDecoder.forward, mp:164-178: blocks, then head(norm(.)) on the LAST return_token_num
    tokens (all tokens if <= 0).  sd: the decoder's own state dict (blocks.*, norm.*, head.*)."""
    D = sd["norm.weight"].shape[0]
    for i in range(_depth(sd)):
        x = block(sd, i, x, heads)
    if return_token_num > 0:
        x = x[:, -return_token_num:]                                   # mp:174
    return F.linear(F.layer_norm(x, (D,), sd["norm.weight"], sd["norm.bias"], LN_EPS), sd["head.weight"], sd["head.bias"])


@torch.no_grad()
def pretrain_forward(sd, x, mask, enc_heads, dec_heads):
    """PretrainVisionTransformer.forward, mp:276-291.  sd: full state dict (encoder.*, decoder.*,
    encoder_to_decoder.weight, mask_token).  Returns [B, N_mask, 1536]: predicted pixels of the masked tokens."""
    x_vis = encoder_forward(_sub(sd, "encoder."), x, mask, enc_heads)  # mp:278  [B, N_vis, C_e]
    x_vis = F.linear(x_vis, sd["encoder_to_decoder.weight"])          # mp:281  (no bias, mp:253)
    B, _, C = x_vis.shape
    pos = sinusoid_table(mask.shape[1], C).expand(B, -1, -1)           # mp:257, mp:285
    pos_vis = pos[~mask].reshape(B, -1, C)                             # mp:286
    pos_mask = pos[mask].reshape(B, -1, C)                             # mp:287
    x_full = torch.cat([x_vis + pos_vis, sd["mask_token"] + pos_mask], dim=1)  # mp:288
    return decoder_forward(_sub(sd, "decoder."), x_full, pos_mask.shape[1], dec_heads)  # mp:289


class TubeMaskingGenerator:
    """masking_generator.py:3-23 restated: one shuffled per-frame mask of int(ratio * H*W) ones, tiled over the
    temporal slots; drawn with np.random exactly as the reference does (same draws for the same np.random.seed)."""

    def __init__(self, input_size, mask_ratio):
        self.frames, self.height, self.width = input_size
        self.num_patches_per_frame = self.height * self.width
        self.total_patches = self.frames * self.num_patches_per_frame
        self.num_masks_per_frame = int(mask_ratio * self.num_patches_per_frame)
        self.total_masks = self.frames * self.num_masks_per_frame

    def __call__(self):
        per_frame = np.hstack([np.zeros(self.num_patches_per_frame - self.num_masks_per_frame),
                               np.ones(self.num_masks_per_frame)])
        np.random.shuffle(per_frame)
        return np.tile(per_frame, (self.frames, 1)).flatten()


def visible_indices(mask):
    """Row-major indices of the surviving tokens of each clip (what x[~mask] keeps, mp:98): int32 [B, n_vis]."""
    B = mask.shape[0]
    idx = [(~mask[b]).nonzero().flatten() for b in range(B)]
    n = {len(i) for i in idx}
    assert len(n) == 1, "every clip must keep the same number of tokens (tube masking guarantees it)"
    return torch.stack(idx).to(torch.int32)
