"""Deterministic synthetic weights and inputs for benchmarks, tools and tests (no arithmetic of the path lives here; the
oracle package re-exports this module as `oracle.synth`).

Everything is generated on the CPU from torch.Generator seeds so the container that writes the golden fixtures and
the GPU box that checks them build bit-identical tensors without shipping hundreds of MB of weights.

Parameter names / shapes follow the reference checkpoint contract (modeling_finetune.py:42-44, 69-78,
143-151, 181-183, 270-272; modeling_pretrain.py:59).
"""
import math

import torch

ARCHS = {
    # name: (embed_dim, depth, heads)         factories: modeling_finetune.py:338-371
    "vit_small_patch16_224": (384, 12, 6),
    "vit_base_patch16_224": (768, 12, 12),
    "vit_large_patch16_224": (1024, 24, 16),
    # reduced-depth variants used only to keep CPU tests fast (same widths, same kernels)
    "vit_small_d2": (384, 2, 6),
    "vit_base_d2": (768, 2, 12),
}
DECODERS = {
    # encoder arch -> (decoder_embed_dim, decoder_num_heads)   factories modeling_pretrain.py:293-363
    "vit_small_patch16_224": (192, 3), "vit_small_d2": (192, 3),
    "vit_base_patch16_224": (384, 6), "vit_base_d2": (384, 6),
    "vit_large_patch16_224": (512, 8),
}
PRETRAIN_ARCHS = {
    # encoders of modeling_pretrain.py:293-387
    "pretrain_videomae_small_patch16_224": "vit_small_patch16_224",
    "pretrain_videomae_base_patch16_224": "vit_base_patch16_224",
    "pretrain_videomae_large_patch16_224": "vit_large_patch16_224",
}
IMG, PATCH, FRAMES, TUBELET, CHANS = 224, 16, 16, 2, 3
N_TOKENS = (FRAMES // TUBELET) * (IMG // PATCH) ** 2  # 1568
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _gen(seed):
    return torch.Generator(device="cpu").manual_seed(int(seed))


def _trunc_normal(shape, std, g):
    # trunc_normal_(std=.02) of the reference init (modeling_finetune.py:285-287), cut at +-2 std
    return (torch.randn(shape, generator=g) * std).clamp_(-2 * std, 2 * std)


def make_state_dict(arch, seed=0, num_classes=2, encoder=False, peaky=1.0):
    """Random-init weights of the named architecture with NON-trivial biases / LN affine / q_bias / v_bias and an
    un-scaled head (the reference's init_scale=0.001 makes every probability exactly 0.5: SURVEY §8c pitfall).
    encoder=True: PretrainVisionTransformerEncoder layout (`norm.*`, no head).  peaky>1 sharpens the softmax."""
    D, depth, heads = ARCHS[arch]
    g = _gen(1000 + seed)
    sd = {}
    fan_in = CHANS * TUBELET * PATCH * PATCH
    bound = 1.0 / math.sqrt(fan_in)  # nn.Conv3d default init range
    sd["patch_embed.proj.weight"] = (torch.rand((D, CHANS, TUBELET, PATCH, PATCH), generator=g) * 2 - 1) * bound
    sd["patch_embed.proj.bias"] = (torch.rand((D,), generator=g) * 2 - 1) * bound
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + 0.1 * torch.randn((D,), generator=g)
        sd[p + "norm1.bias"] = 0.05 * torch.randn((D,), generator=g)
        sd[p + "attn.q_bias"] = 0.02 * torch.randn((D,), generator=g)
        sd[p + "attn.v_bias"] = 0.02 * torch.randn((D,), generator=g)
        w = _trunc_normal((3 * D, D), 0.02, g)
        if peaky != 1.0:
            w[: 2 * D] *= peaky  # scale q and k rows: logits grow by peaky^2
        sd[p + "attn.qkv.weight"] = w
        sd[p + "attn.proj.weight"] = _trunc_normal((D, D), 0.02, g)
        sd[p + "attn.proj.bias"] = 0.02 * torch.randn((D,), generator=g)
        sd[p + "norm2.weight"] = 1.0 + 0.1 * torch.randn((D,), generator=g)
        sd[p + "norm2.bias"] = 0.05 * torch.randn((D,), generator=g)
        sd[p + "mlp.fc1.weight"] = _trunc_normal((4 * D, D), 0.02, g)
        sd[p + "mlp.fc1.bias"] = 0.02 * torch.randn((4 * D,), generator=g)
        sd[p + "mlp.fc2.weight"] = _trunc_normal((D, 4 * D), 0.02, g)
        sd[p + "mlp.fc2.bias"] = 0.02 * torch.randn((D,), generator=g)
    if encoder:
        sd["norm.weight"] = 1.0 + 0.1 * torch.randn((D,), generator=g)
        sd["norm.bias"] = 0.05 * torch.randn((D,), generator=g)
    else:
        sd["fc_norm.weight"] = 1.0 + 0.1 * torch.randn((D,), generator=g)
        sd["fc_norm.bias"] = 0.05 * torch.randn((D,), generator=g)
        # head wide enough that the two logits differ by O(1): p is neither 0.5 nor saturated
        sd["head.weight"] = 0.05 * torch.randn((num_classes, D), generator=g)
        sd["head.bias"] = 0.1 * torch.randn((num_classes,), generator=g)
    return sd


OUTLIER_CHANNELS = (7, 123, 300)  # valid for every width in ARCHS


def make_trained_like_state_dict(arch, seed=0, num_classes=2, init_values=0.1):
    """Weights with the statistics of a TRAINED video ViT rather than of trunc-normal init — the regime where a bf16
    residual stream and LayerNorm statistics taken from E[x^2] - mean^2 partial sums could lose accuracy:
      * massive-activation channels: three channels of the residual stream sit 60-100 x above the typical |x| (the
        patch-embed bias puts them there, the fc2 biases of the first blocks keep feeding them);
      * token rows with a mean of several sigma (every channel of the patch-embed bias is shifted);
      * LayerNorm gamma spread log-uniformly over [0.1, 5] and beta ~ 0.3 N(0, 1) in every block;
      * layer scale (init_values > 0): gamma_1 / gamma_2 = init_values (1 + 0.3 N(0, 1)), modeling_finetune.py:153-162.
    The head and fc_norm keep make_state_dict's scale so that the probabilities stay informative."""
    D, depth, _ = ARCHS[arch]
    sd = make_state_dict(arch, seed=seed, num_classes=num_classes)
    g = _gen(9000 + seed)
    ch = torch.tensor(OUTLIER_CHANNELS)
    sign = torch.tensor([1.0, -1.0, 1.0])
    sd["patch_embed.proj.bias"] = sd["patch_embed.proj.bias"] + 1.5
    sd["patch_embed.proj.bias"][ch] += sign * (40.0 + 20.0 * torch.rand((3,), generator=g))
    lo, hi = math.log(0.1), math.log(5.0)
    for i in range(depth):
        p = f"blocks.{i}."
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = torch.exp(lo + (hi - lo) * torch.rand((D,), generator=g))
            sd[p + n + ".bias"] = 0.3 * torch.randn((D,), generator=g)
        if i < 2:
            sd[p + "mlp.fc2.bias"][ch] += sign * (100.0 + 50.0 * torch.rand((3,), generator=g))
        sd[p + "gamma_1"] = init_values * (1.0 + 0.3 * torch.randn((D,), generator=g))
        sd[p + "gamma_2"] = init_values * (1.0 + 0.3 * torch.randn((D,), generator=g))
    return sd


def make_variant_state_dict(arch, seed=0, num_classes=2, tubelet=TUBELET, final_reduction="fc_norm", cls_token=False):
    """make_state_dict re-keyed for the other forms of the classifier: final_reduction 'cls' / 'none' own `norm.*`
    instead of `fc_norm.*` (modeling_finetune.py:269-270); tubelet 1 (the UMT job) keeps the first temporal slice of
    the Conv3d weight; cls_token adds MVD's `cls_token` [1, 1, D] (other_models/MVD/modeling_finetune.py:364-366)."""
    D, _, _ = ARCHS[arch]
    sd = make_state_dict(arch, seed=seed, num_classes=num_classes)
    if tubelet != TUBELET:
        assert tubelet == 1
        sd["patch_embed.proj.weight"] = sd["patch_embed.proj.weight"][:, :, :1].contiguous() * math.sqrt(2.0)
    if final_reduction != "fc_norm":
        sd["norm.weight"] = sd.pop("fc_norm.weight")
        sd["norm.bias"] = sd.pop("fc_norm.bias")
    if cls_token:
        sd["cls_token"] = 0.5 * torch.randn((1, 1, D), generator=_gen(7000 + seed))
    return sd


def _block_weights(sd, p, D, g):
    sd[p + "norm1.weight"] = 1.0 + 0.1 * torch.randn((D,), generator=g)
    sd[p + "norm1.bias"] = 0.05 * torch.randn((D,), generator=g)
    sd[p + "attn.q_bias"] = 0.02 * torch.randn((D,), generator=g)
    sd[p + "attn.v_bias"] = 0.02 * torch.randn((D,), generator=g)
    sd[p + "attn.qkv.weight"] = _trunc_normal((3 * D, D), 0.04, g)
    sd[p + "attn.proj.weight"] = _trunc_normal((D, D), 0.04, g)
    sd[p + "attn.proj.bias"] = 0.02 * torch.randn((D,), generator=g)
    sd[p + "norm2.weight"] = 1.0 + 0.1 * torch.randn((D,), generator=g)
    sd[p + "norm2.bias"] = 0.05 * torch.randn((D,), generator=g)
    sd[p + "mlp.fc1.weight"] = _trunc_normal((4 * D, D), 0.04, g)
    sd[p + "mlp.fc1.bias"] = 0.02 * torch.randn((4 * D,), generator=g)
    sd[p + "mlp.fc2.weight"] = _trunc_normal((D, 4 * D), 0.04, g)
    sd[p + "mlp.fc2.bias"] = 0.02 * torch.randn((D,), generator=g)


def make_pretrain_state_dict(arch, seed=0, decoder_depth=4):
    """State dict of PretrainVisionTransformer (modeling_pretrain.py:183-258) for the encoder `arch`:
    encoder.* (as make_state_dict(encoder=True)), decoder.{blocks.*, norm.*, head.*}, encoder_to_decoder.weight,
    mask_token — all non-trivial so every bias / affine path is exercised."""
    D, _, _ = ARCHS[arch]
    Dd, _ = DECODERS[arch]
    sd = {"encoder." + k: v for k, v in make_state_dict(arch, seed=seed, encoder=True).items()}
    g = _gen(5000 + seed)
    for i in range(decoder_depth):
        _block_weights(sd, f"decoder.blocks.{i}.", Dd, g)
    sd["decoder.norm.weight"] = 1.0 + 0.1 * torch.randn((Dd,), generator=g)
    sd["decoder.norm.bias"] = 0.05 * torch.randn((Dd,), generator=g)
    n_pix = CHANS * TUBELET * PATCH * PATCH
    sd["decoder.head.weight"] = _trunc_normal((n_pix, Dd), 0.05, g)
    sd["decoder.head.bias"] = 0.05 * torch.randn((n_pix,), generator=g)
    sd["encoder_to_decoder.weight"] = _trunc_normal((Dd, D), 0.04, g)
    sd["mask_token"] = 0.5 * torch.randn((1, 1, Dd), generator=g)
    return sd


def bf16_round(x):
    """Both arms of every parity test see the SAME values: fp32 numbers that are exactly representable in bf16."""
    return x.to(torch.bfloat16).to(torch.float32)


def make_clips(B, seed=0, frames=FRAMES, img=IMG):
    """x ~ N(0,1) [B,3,16,224,224] as in test_efficiency.py:17, rounded to bf16-representable fp32."""
    return bf16_round(torch.randn((B, CHANS, frames, img, img), generator=_gen(2000 + seed)))


def make_video(T, seed=0):
    """DoTA-shaped synthetic video: T frames of smooth-ish uint8 noise, ImageNet-normalised like prepare_image
    (run_inference.py:15-34) -> frames [T,3,224,224] fp32 (bf16-representable)."""
    g = _gen(3000 + seed)
    base = torch.rand((1, CHANS, IMG, IMG), generator=g)
    drift = torch.rand((T, CHANS, IMG, IMG), generator=g)
    mix = torch.linspace(0.15, 0.6, T).view(T, 1, 1, 1)
    u8 = ((1 - mix) * base + mix * drift).mul(255).round().clamp(0, 255)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return bf16_round((u8 / 255.0 - mean) / std)


def windows_from_video(frames, start=0, count=None, stride=1):
    """Sliding windows of 16 consecutive frames, stride 1 (dota.py:204-223, sequencing.py:38-62):
    frames [T,3,H,W] -> clips [n,3,16,H,W]; window w covers frames [w, w+16), label = its last frame."""
    T = frames.shape[0]
    n_all = (T - FRAMES) // stride + 1
    count = n_all - start if count is None else count
    idx = [start + i for i in range(count)]
    return torch.stack([frames[w * stride: w * stride + FRAMES].permute(1, 0, 2, 3) for w in idx]).contiguous()


def tube_mask(B, ratio=0.9, seed=0):
    """TubeMaskingGenerator (masking_generator.py:3-23): one random per-frame mask of int(ratio*196) ones tiled over
    the 8 temporal slots, one independent mask per clip.  Uses a seeded torch permutation instead of np.random.shuffle
    (any permutation is a valid draw).  Returns bool [B,1568], True = masked."""
    per_frame = (IMG // PATCH) ** 2
    n_mask = int(ratio * per_frame)
    g = _gen(4000 + seed)
    out = torch.zeros((B, FRAMES // TUBELET, per_frame), dtype=torch.bool)
    for b in range(B):
        perm = torch.randperm(per_frame, generator=g)
        out[b, :, perm[:n_mask]] = True
    return out.reshape(B, -1)
