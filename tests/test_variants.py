"""The other forms of the classifier: final_reduction 'cls' / 'none' of modeling_finetune.py and the
MVD / UMT sibling architectures (other_models/ in the reference).

CPU: the oracle against the golden outputs of the unmodified reference modules, and this package's host-side position
tables / module structure against the same fixtures.  GPU (-m gpu): the drop-in modules through the C ABI against the
golden outputs, with BASELINE.json's tolerances."""
import numpy as np
import pytest
import torch

from oracle import vit_oracle
from oracle.make_golden import VARIANTS, variant_setup
from tests import parity

FAST = [n for n in VARIANTS if n != "var_mvd_vitb_b4"]


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_reference_variants(name):
    g = parity.golden(name)
    family, arch, extra, sd, x, pos, red = variant_setup(name)
    heads = parity.synth.ARCHS[arch][2]
    logits, feat = vit_oracle.vit_forward_variant(sd, x, heads, pos, final_reduction=red, mvd=family == "mvd")
    assert logits.shape == g["logits"].shape
    assert float((logits - torch.from_numpy(g["logits"])).abs().max()) <= 2e-5
    assert np.allclose(feat.norm(dim=-1).numpy(), g["feat_norm"], rtol=1e-5)
    tok = torch.from_numpy(g["pos_tok"])
    ch = torch.from_numpy(g["hid_ch"])
    assert np.array_equal(pos[0][tok][:, ch].numpy(), g["pos_samples"])


@pytest.mark.parametrize("name", FAST)
def test_package_position_tables_and_structure(name):
    """Host side of the drop-in modules: position table bit-identical to the reference's, state-dict keys as the
    reference's (strict load succeeded in build_variant), attributes the runners read."""
    g = parity.golden(name)
    model, x, info = parity.build_variant(name, device=None)
    pos = model.pos_embed.detach()
    assert pos.shape == info["pos"].shape and pos.dtype == torch.float32
    assert torch.equal(pos, info["pos"]), "package position table differs from the oracle's"
    tok = torch.from_numpy(g["pos_tok"])
    ch = torch.from_numpy(g["hid_ch"])
    assert np.array_equal(pos[0][tok][:, ch].numpy(), g["pos_samples"])
    assert np.allclose([pos.double().sum().item(), pos.double().abs().sum().item()], g["pos_sum"], rtol=1e-12, atol=1e-9)
    keys = set(model.state_dict())
    assert ("cls_token" in keys) == bool(VARIANTS[name][4].get("use_cls_token"))
    assert ("norm.weight" in keys) == (info["red"] != "fc_norm") and ("fc_norm.weight" in keys) == (info["red"] == "fc_norm")
    assert isinstance(model.pos_embed, torch.nn.Parameter) == ("pos_embed" in keys)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(x[:1])


def test_sibling_factories_registered():
    from simple_tad_b200.other_models.MVD import modeling_finetune as mvd
    from simple_tad_b200.other_models.UMT import modeling_finetune as umt
    from simple_tad_b200.registry import create_model, is_model
    for n in ("mvd_vit_small_patch16_224", "mvd_vit_base_patch16_224", "mvd_vit_large_patch16_224",
              "mvd_vit_huge_patch16_224", "umt_vit_base_patch16_224", "umt_vit_large_patch16_384"):
        assert is_model(n)
    m = create_model("mvd_vit_small_patch16_224", pretrained=False, num_classes=2, all_frames=16, tubelet_size=2,
                     drop_block_rate=None, use_cls_token=True, final_reduction="fc_norm", init_scale=0.001)
    assert m.cls_token.shape == (1, 1, 384) and m.pos_embed.shape == (1, 1568, 384) and m.patch_size == 16
    assert (m.patch_embed.num_patches_t, m.patch_embed.num_patches_h, m.patch_embed.num_patches_w) == (8, 14, 14)
    assert m.no_weight_decay() == {"pos_embed", "cls_token"} and m.get_num_layers() == 12
    u = umt.vit_small_patch16_224(num_classes=2, all_frames=8, tubelet_size=1)
    assert u.pos_embed.shape == (1, 1568, 384) and not isinstance(u.pos_embed, torch.nn.Parameter)
    assert mvd.get_sinusoid_encoding_table(8, 16, cls_token=True).shape == (1, 8, 16)
    assert mvd.get_2d_sincos_pos_embed(16, 3, cls_token=True).shape == (10, 16)


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
def test_variants_vs_reference_gpu(name):
    g = parity.golden(name)
    model, x, info = parity.build_variant(name, device="cuda")
    xd = x.to("cuda")
    logits = model(xd)
    assert tuple(logits.shape) == g["logits"].shape
    if logits.dim() == 2:
        parity.check_logits(logits, g["logits"], name)
    else:
        # per-token logits ('none'): 2-vectors near the origin make a per-row cosine meaningless; the probability
        # bound holds for every token and the cosine is taken over all of them
        ref = torch.from_numpy(g["logits"]).double()
        got = logits.double().cpu()
        dp = float((got.softmax(-1) - ref.softmax(-1)).abs().max())
        assert dp <= parity.TOL_DP, f"{name}: max|dp| over tokens = {dp:.3e}"
        assert parity.logit_cosine(got, ref) >= parity.TOL_COS
    feat = model.forward_features(xd).float().cpu()
    ch = torch.from_numpy(g["hid_ch"])
    assert np.allclose(feat.norm(dim=-1).numpy(), g["feat_norm"], rtol=2e-2)
    feat_s = feat[..., ch] if feat.dim() == 2 else feat[:, ::97][..., ch]
    ref_s = torch.from_numpy(g["feat_samples"])
    rel = float((feat_s - ref_s).norm() / ref_s.norm())
    assert rel <= parity.TOL_HIDDEN_REL_L2, f"{name}: feature samples rel-L2 {rel:.3e}"
    if info["red"] == "none" and info["family"] != "mvd":
        logits2, probs = model.forward_probs(xd)
        assert probs.shape == logits.shape
        assert float((probs.cpu() - torch.from_numpy(g["probs"])).abs().max()) <= parity.TOL_DP


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["var_mvd_vits_d2_b2", "var_mvd_clstok_vits_d2_b2"])
def test_mvd_windows_equal_clips_gpu(name):
    """MVD (with and without its class token) through the sliding-window frame path (the MVD-B DoTA job evaluates
    frame-level): windows read out of the frame buffer == materialised clips, bit for bit."""
    model, x, info = parity.build_variant(name, device="cuda")
    frames = parity.synth.make_video(20, seed=3).to("cuda")
    lw, pw = model.forward_windows(frames, reuse_tubelets=False)
    lw = lw.clone()
    clips = parity.synth.windows_from_video(frames.cpu()).to("cuda")
    lc = model(clips).clone()
    assert torch.equal(lw, lc)
    # default: tubelet embeddings shared between the overlapping windows (not with a class token) — one extra bf16 rounding
    lr, _ = model.forward_windows(frames)
    assert float((lr - lc).abs().max()) <= 5e-3


@pytest.mark.gpu
def test_frame_step_windows_follow_the_sequencer_gpu():
    """A 30 fps video scored at 10 fps (DADA-2000, dada.py:31,173-177): the windows the kernel reads with
    frame_step = 3 are the index lists of RegularSequencer (tests/golden/sequencer.npz), and the runner scores exactly
    those, aligned to the last frame."""
    from simple_tad_b200 import sequencing
    from simple_tad_b200.runner import SlidingWindowRunner
    model, _, _ = parity.build_variant("var_mvd_vits_d2_b2", device="cuda")
    T_video, step = 58, 5
    frames = parity.synth.make_video(T_video, seed=5)
    plan = sequencing.window_plan(T_video, input_frequency=30, seq_frequency=10, seq_length=16, step=step)
    assert plan.frame_step == 3 and plan.count == 3 and plan.start == 2
    idx = torch.tensor(plan.sequences())                                        # [count, 16] source frame indices
    clips = frames[idx].permute(0, 2, 1, 3, 4).contiguous().to("cuda")          # [count, C, T, H, W]
    ref = model(clips)
    got, _ = model.forward_windows(frames.to("cuda"), start=plan.start, count=plan.count, stride=plan.stride,
                                   frame_step=plan.frame_step)
    assert torch.equal(got, ref)
    runner = SlidingWindowRunner(model, batch_windows=8, stride=step, frame_step=3)   # one batch: same tiles as `ref`
    lg, _ = runner.score_frames(frames)
    assert torch.equal(lg, ref.cpu())
    lg_all = runner.score_videos([frames, frames[:50]])
    plan2 = sequencing.window_plan(50, 30, 10, 16, step)
    # the runner batches windows across video boundaries (one batch of 3 + 2 windows here): another batch size than
    # `ref`, so equal within the parity tolerance rather than bit for bit
    assert lg_all.shape[0] == plan.count + plan2.count
    parity.check_logits(lg_all[: plan.count], ref, "frame-step windows, two videos in one batch")
    # batches of two windows: other tile shapes, so equal within the parity tolerance rather than bit for bit
    lg2, _ = SlidingWindowRunner(model, batch_windows=2, stride=step, frame_step=3).score_frames(frames)
    parity.check_logits(lg2, ref, "frame-step windows, batches of 2")
