"""Shared helpers of the parity tests: golden fixtures, metrics, model builders."""
import os

import numpy as np
import torch

from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json / BASELINE.md §6: per-frame anomaly probability max |dp| <= 1e-2, logit cosine >= 0.999
TOL_DP = 1e-2
TOL_COS = 0.999
# hidden states (bf16 residual stream vs fp32 oracle): relative L2 per block (<= ~2e-2)
TOL_HIDDEN_REL_L2 = 2e-2


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def logit_cosine(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def row_cosine_min(a, b):
    """Minimum over rows of the cosine between matching rows (per-clip logits, or per-token features)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    a = a.reshape(-1, a.shape[-1])
    b = b.reshape(-1, b.shape[-1])
    cos = (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1) + 1e-300)
    return float(cos.min())


def check_logits(got, ref, what):
    got = torch.as_tensor(got).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    dp = float((got.softmax(-1) - ref.softmax(-1)).abs().max())
    cos_all = logit_cosine(got, ref)
    cos_row = row_cosine_min(got, ref)
    assert torch.isfinite(got).all(), f"{what}: non-finite logits"
    assert dp <= TOL_DP, f"{what}: max|dp| = {dp:.3e} > {TOL_DP} (cos {cos_all:.6f})"
    assert cos_all >= TOL_COS, f"{what}: logit cosine {cos_all:.6f} < {TOL_COS} (max|dp| {dp:.3e})"
    assert cos_row >= TOL_COS, f"{what}: worst per-clip logit cosine {cos_row:.6f} < {TOL_COS}"
    return {"what": what, "max_dp": dp, "cos": cos_all, "cos_row_min": cos_row}


def build_classifier(arch, sd, device="cuda"):
    """The drop-in model of this repo with the reference-format state dict loaded."""
    from simple_tad_b200 import modeling_finetune as mf
    from functools import partial
    D, depth, heads = synth.ARCHS[arch]
    model = mf.VisionTransformer(patch_size=16, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16,
                                 tubelet_size=2, init_scale=1.0, final_reduction="fc_norm")
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def build_encoder(arch, sd, device="cuda"):
    from simple_tad_b200 import modeling_pretrain as mp
    from functools import partial
    D, depth, heads = synth.ARCHS[arch]
    model = mp.PretrainVisionTransformerEncoder(embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                                norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def build_pretrain(arch, sd, decoder_depth, device="cuda"):
    """The full PretrainVisionTransformer drop-in (encoder + decoder) with a reference-format state dict loaded."""
    from simple_tad_b200 import modeling_pretrain as mp
    from functools import partial
    D, depth, heads = synth.ARCHS[arch]
    Dd, dheads = synth.DECODERS[arch]
    model = mp.PretrainVisionTransformer(
        img_size=224, patch_size=16, encoder_embed_dim=D, encoder_depth=depth, encoder_num_heads=heads,
        encoder_num_classes=0, decoder_num_classes=1536, decoder_embed_dim=Dd, decoder_num_heads=dheads,
        decoder_depth=decoder_depth, mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def build_variant(name, device="cuda"):
    """The drop-in model of a VARIANTS fixture (final_reduction 'cls' / 'none', the MVD / UMT siblings) with the
    reference-format state dict loaded.  Returns (model, clips, oracle pieces)."""
    from functools import partial
    from oracle.make_golden import variant_setup
    family, arch, extra, sd, x, pos, red = variant_setup(name)
    if family == "mvd":
        from simple_tad_b200.other_models.MVD import modeling_finetune as mod
    elif family == "umt":
        from simple_tad_b200.other_models.UMT import modeling_finetune as mod
    else:
        from simple_tad_b200 import modeling_finetune as mod
    D, depth, heads = synth.ARCHS[arch]
    kw = dict(patch_size=16, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16, tubelet_size=2,
              init_scale=1.0, final_reduction="fc_norm")
    kw.update(extra)
    model = mod.VisionTransformer(**kw)
    sd_load = dict(sd)
    if isinstance(model.pos_embed, torch.nn.Parameter):  # UMT: an interpolated table is a parameter (umt:237-239)
        sd_load["pos_embed"] = model.pos_embed.detach().clone()
    model.load_state_dict(sd_load, strict=True)
    if device is not None:
        model = model.to(device)
    return model.eval(), x, dict(family=family, arch=arch, sd=sd, pos=pos, red=red, heads=heads)
