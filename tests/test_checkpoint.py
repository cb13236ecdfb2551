"""Host logic of checkpoint loading (run_frame_finetuning.py:396-460, utils.py:336-383): a DAPT / VideoMAE pre-training
checkpoint or a released fine-tuned one becomes the classifier's state dict.  CPU only."""
from functools import partial

import torch

from oracle import synth
from simple_tad_b200 import checkpoint as ck, modeling_finetune as mf

ARCH = "vit_small_d2"


def _classifier(**kw):
    D, depth, heads = synth.ARCHS[ARCH]
    args = dict(patch_size=16, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16, tubelet_size=2,
                init_scale=1.0)
    args.update(kw)
    return mf.VisionTransformer(**args)


def test_dapt_checkpoint_feeds_the_classifier():
    """encoder.* -> *, encoder.norm.* -> fc_norm.*, decoder / mask token reported as unused, head left at its init."""
    sd = synth.make_pretrain_state_dict(ARCH, seed=3, decoder_depth=1)
    model = _classifier()
    head0 = model.head.weight.detach().clone()
    missing = ck.load_finetune_checkpoint(model, {"model": dict(sd)}, verbose=False)
    assert sorted(missing) == ["head.bias", "head.weight"]
    own = model.state_dict()
    for k, v in sd.items():
        if k.startswith("encoder.norm."):
            assert torch.equal(own[k.replace("encoder.norm", "fc_norm")], v)
        elif k.startswith("encoder."):
            assert torch.equal(own[k[8:]], v), k
    assert torch.equal(model.head.weight, head0)
    assert not any(k.startswith("decoder") or k == "mask_token" for k in own)


def test_select_key_backbone_prefix_and_head_of_another_width():
    sd = synth.make_state_dict(ARCH, seed=4, num_classes=400)          # e.g. a Kinetics-400 fine-tuned checkpoint
    wrapped = {"module": {"backbone." + k: v for k, v in sd.items()}}
    assert ck.select_state_dict(wrapped) is wrapped["module"]
    assert ck.select_state_dict({"x": 1}) == {"x": 1}
    model = _classifier()
    remapped = ck.remap_finetune_keys(dict(wrapped["module"]), model)
    assert "blocks.0.attn.qkv.weight" in remapped and not any(k.startswith("backbone.") for k in remapped)
    # the head check of rff:413-417 runs BEFORE the prefixes are stripped: a `backbone.head.*` of another width survives
    # the removal and is then reported as a size mismatch by the loader instead of being loaded
    missing = ck.load_state_dict(model, remapped, verbose=False)
    assert sorted(missing) == ["head.bias", "head.weight"]
    assert torch.equal(model.state_dict()["blocks.1.mlp.fc2.weight"], sd["blocks.1.mlp.fc2.weight"])
    # un-prefixed checkpoint: the 400-way head is dropped from the checkpoint dict itself
    plain = dict(sd)
    out = ck.remap_finetune_keys(plain, model)
    assert "head.weight" not in plain and "head.weight" not in out and "fc_norm.weight" in out


def test_prefix_and_ignore_missing():
    sd = synth.make_state_dict(ARCH, seed=5)
    model = _classifier()
    pref = {"net." + k: v for k, v in sd.items() if not k.startswith("fc_norm")}
    pref["other.thing"] = torch.zeros(1)
    missing = ck.load_state_dict(model, pref, prefix="net.", ignore_missing="fc_norm.bias", verbose=False)
    assert missing == ["fc_norm.weight"]
    assert torch.equal(model.state_dict()["head.weight"], sd["head.weight"])


def test_pos_embed_interpolation_matches_the_reference_recipe():
    """rff:432-458 on a learnable table: 8 x 14 x 14 rows of a 224 px checkpoint -> 8 x 20 x 20 rows of a 320 px model."""
    D = synth.ARCHS[ARCH][0]
    model = _classifier(img_size=320, use_learnable_pos_emb=True)
    assert model.pos_embed.shape == (1, 8 * 20 * 20, D)
    g = torch.Generator().manual_seed(0)
    pos = torch.randn(1, 8 * 14 * 14, D, generator=g)
    sd = {"pos_embed": pos.clone()}
    ck.interpolate_pos_embed(sd, model, num_frames=16)
    ref = pos.reshape(8, 14, 14, D).permute(0, 3, 1, 2)
    ref = torch.nn.functional.interpolate(ref, size=(20, 20), mode="bicubic", align_corners=False)
    ref = ref.permute(0, 2, 3, 1).reshape(1, 8 * 400, D)
    assert sd["pos_embed"].shape == (1, 3200, D) and torch.equal(sd["pos_embed"], ref)
    # same grid: untouched
    same = {"pos_embed": torch.randn(1, 3200, D, generator=g)}
    before = same["pos_embed"].clone()
    ck.interpolate_pos_embed(same, model, num_frames=16)
    assert torch.equal(same["pos_embed"], before)
    missing = ck.load_state_dict(model, sd, verbose=False)
    assert "pos_embed" not in missing and torch.equal(model.pos_embed.detach(), ref)
