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
    """mf:185-191: Conv3d(k = s = (tubelet,16,16)) then flatten(2).transpose(1,2): tokens ordered (t', h', w')."""
    w = sd["patch_embed.proj.weight"]
    y = F.conv3d(x, w, sd["patch_embed.proj.bias"], stride=(w.shape[2], synth.PATCH, synth.PATCH))
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
    """Block.forward, mf:159-166: without layer scale (gamma_1 None, init_values = 0) or with it (mf:164-165)."""
    p = f"blocks.{i}."
    D = x.shape[-1]
    a = attention(sd, p, F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS), heads)
    x = x + (sd[p + "gamma_1"] * a if p + "gamma_1" in sd else a)
    m = mlp(sd, p, F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS))
    return x + (sd[p + "gamma_2"] * m if p + "gamma_2" in sd else m)


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


def _sincos_1d(dim, pos, scale=None):
    """mvd:102-122 (other_models/MVD/modeling_finetune.py): [M, dim] = sin | cos of pos x 10000^(-k / (dim/2)), float64."""
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=float) / (dim / 2.0))
    pos = np.asarray(pos).reshape(-1)
    if scale is not None:
        pos = pos * scale
    ang = pos[:, None] * omega[None, :]
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_3d_table(d_hid, grid_size, t_size):
    """MVD position table, mvd:24-69: per token (t', h', w'): [ 1-D sincos of t' (D/4) | 1-D sincos of w' (3D/8) | 1-D
    sincos of h' (3D/8) ] — np.meshgrid(w, h) puts the w' coordinate first (mvd:38-39, mvd:92-97).  [1, T*H*W, D] fp32."""
    d_t, d_s = d_hid // 4, d_hid // 4 * 3
    ax = np.arange(grid_size, dtype=np.float32)
    ww = np.tile(ax[None, :], (grid_size, 1)).reshape(-1)    # w' of token (h', w') in row-major order
    hh = np.tile(ax[:, None], (1, grid_size)).reshape(-1)    # h'
    spatial = np.concatenate([_sincos_1d(d_s // 2, ww), _sincos_1d(d_s // 2, hh)], axis=1)    # [H*W, 3D/4]
    temporal = _sincos_1d(d_t, np.arange(t_size, dtype=np.float32))                           # [T, D/4]
    tab = np.concatenate([np.repeat(temporal[:, None], grid_size ** 2, 1), np.repeat(spatial[None], t_size, 0)], -1)
    return torch.FloatTensor(tab.reshape(-1, d_hid)).unsqueeze(0)


def umt_table(n_position, d_hid, cur_frame, pre_n_position=1568):
    """UMT position table, umt:195-239 (other_models/UMT/modeling_finetune.py): the 1568-row sinusoid table of the 8 x 14
    x 14 pre-training grid, bicubic over (h', w') if the image grid differs, linear over t' if the clip length does."""
    tab = sinusoid_table(pre_n_position, d_hid)
    if cur_frame != -1 and n_position // cur_frame * 8 != pre_n_position:     # umt:208-221
        new_p = int((n_position // cur_frame) ** 0.5)
        t = tab.reshape(8, 14, 14, d_hid).permute(0, 3, 1, 2)
        t = F.interpolate(t, size=(new_p, new_p), mode="bicubic", align_corners=False)
        tab = t.permute(0, 2, 3, 1).reshape(1, 8 * new_p * new_p, d_hid)
    if cur_frame != -1 and cur_frame != 8:                                    # umt:222-234
        p = int((n_position // cur_frame) ** 0.5)
        t = tab.reshape(8, p * p, d_hid).permute(1, 2, 0)                     # [HW, C, T]
        t = F.interpolate(t, size=cur_frame, mode="linear")
        tab = t.permute(2, 0, 1).reshape(1, cur_frame * p * p, d_hid)
    return tab


@torch.no_grad()
def vit_forward_variant(sd, x, heads, pos, final_reduction="fc_norm", mvd=False):
    """The classifier forward in its other forms: any position table `pos` [1, N, D] (mf:312-313), MVD's class token
    prepended after the position add when `cls_token` is in sd (mvd:428-435), and the three reductions —
    'fc_norm': fc_norm(mean of the patch tokens) (mf:325-326; mvd:447-449 drops the class token first),
    'cls': norm(x)[:, 0] (mf:327-328), 'none': norm(x) per token (mf:329-330; MVD returns x[:, 0] here too, mvd:450-451)
    — followed by the head (mf:334).  Returns (logits, features)."""
    D = sd["patch_embed.proj.bias"].shape[0]
    h = patch_embed(sd, x) + pos
    if "cls_token" in sd:
        h = torch.cat((sd["cls_token"].expand(h.shape[0], -1, -1), h), dim=1)
    for i in range(_depth(sd)):
        h = block(sd, i, h, heads)
    if final_reduction == "fc_norm":
        if "cls_token" in sd:
            h = h[:, 1:]
        feat = F.layer_norm(h.mean(1), (D,), sd["fc_norm.weight"], sd["fc_norm.bias"], LN_EPS)
    else:
        h = F.layer_norm(h, (D,), sd["norm.weight"], sd["norm.bias"], LN_EPS)
        feat = h[:, 0] if (final_reduction == "cls" or mvd) else h
    return F.linear(feat, sd["head.weight"], sd["head.bias"]), feat


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
