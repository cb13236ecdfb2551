"""Whole-path parity on the GPU: the drop-in modules (C ABI -> sm_100a kernels) against
  (1) the golden outputs of the UNMODIFIED reference (tests/golden/, oracle/make_golden.py) and
  (2) the CPU oracle run live on the same seeded inputs (hidden states per block, module-level checks).
Tolerances are BASELINE.json's: max|dp| <= 1e-2, logit cosine >= 0.999 (tests/parity.py)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import synth, vit_oracle
from tests import parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _hidden_states_gpu(model, x):
    """Residual stream after patch-embed(+pos) and after every Block, through the module-level drop-in path."""
    from simple_tad_b200 import _lib
    prep = model.prepare(x.device)
    B = x.shape[0]
    h = model.patch_embed(x) + model.pos_embed.to(x.device)
    hs = [h]
    for blk in model.blocks:
        h = blk(h)
        hs.append(h)
    return hs


@pytest.mark.parametrize("fixture,arch,seed,peaky", [
    ("small_vits_d2_b2", "vit_small_d2", 11, 1.0),
    ("peaky_vits_d2_b2", "vit_small_d2", 13, 3.0),
])
def test_small_models_vs_reference_and_oracle_hidden(fixture, arch, seed, peaky):
    g = parity.golden(fixture)
    sd = synth.make_state_dict(arch, seed=seed, peaky=peaky)
    x = synth.make_clips(2, seed=seed)
    model = parity.build_classifier(arch, sd)
    logits = model(x.to(DEV))
    parity.check_logits(logits, g["logits"], fixture)
    # hidden states: module-level path vs the oracle's fp32 residual stream, block by block
    D, depth, heads = synth.ARCHS[arch]
    _, hid_ref = vit_oracle.vit_forward(sd, x, heads, return_hidden=True)
    hid = _hidden_states_gpu(model, x.to(DEV))
    assert len(hid) == len(hid_ref)
    for i, (a, b) in enumerate(zip(hid, hid_ref)):
        rel = float((a.float().cpu() - b).norm() / b.norm())
        assert rel <= parity.TOL_HIDDEN_REL_L2, f"{fixture}: hidden state {i} rel-L2 {rel:.3e}"


def test_config1_vits_batch4():
    """BASELINE config 1: ViT-S/16, batch-4 synthetic clips, probabilities of run_inference_simple's model."""
    g = parity.golden("c1_vits_b4")
    sd = synth.make_state_dict("vit_small_patch16_224", seed=1)
    model = parity.build_classifier("vit_small_patch16_224", sd)
    logits, probs = model.forward_probs(synth.make_clips(4, seed=1).to(DEV))
    parity.check_logits(logits, g["logits"], "config1 logits")
    dp = float((probs.cpu() - torch.from_numpy(g["probs_ris"])).abs().max())
    assert dp <= parity.TOL_DP, f"config1: max|dp| vs run_inference_simple = {dp:.3e}"
    feats = model.forward_features(synth.make_clips(4, seed=1).to(DEV))
    assert feats.shape == (4, 384) and torch.isfinite(feats).all()


def test_config2_vitb_sliding_window_video():
    """BASELINE config 2: ViT-B/16, one DoTA-shaped 100-frame video -> 85 stride-1 windows, per-frame scores.
    The windows are read straight out of the resident frame buffer (no clip materialisation)."""
    g = parity.golden("c2_vitb_video100")
    sd = synth.make_state_dict("vit_base_patch16_224", seed=2)
    model = parity.build_classifier("vit_base_patch16_224", sd)
    frames = synth.make_video(100, seed=2).to(DEV)
    logits, probs = model.forward_windows(frames)
    assert logits.shape == (85, 2)
    parity.check_logits(logits, g["logits"], "config2 (85 windows)")
    # same windows as explicit clips: identical kernels, identical tiles -> bit-identical scores
    clips = synth.windows_from_video(synth.make_video(100, seed=2), start=0, count=85).to(DEV)
    logits2 = model(clips)
    assert torch.equal(logits, logits2), "frame-buffer windows and materialised clips disagree"


def test_config3_vitl_two_videos():
    """BASELINE config 3 (single-GPU part): ViT-L/16 sliding windows (2 videos x 20 frames -> 10 windows)."""
    g = parity.golden("c3_vitl_2x20")
    sd = synth.make_state_dict("vit_large_patch16_224", seed=3)
    model = parity.build_classifier("vit_large_patch16_224", sd)
    outs = []
    for v in range(2):
        frames = synth.make_video(20, seed=3 + v).to(DEV)
        outs.append(model.forward_windows(frames)[0])
    parity.check_logits(torch.cat(outs), g["logits"], "config3 (ViT-L, 10 windows)")


def test_config4_masked_encoder():
    """BASELINE config 4: ViT-B encoder, 90 % tube masking -> [B, 160, 768]; cosine >= 0.999 per token row."""
    g = parity.golden("c4_enc_vitb_b4")
    sd = synth.make_state_dict("vit_base_patch16_224", seed=4, encoder=True)
    model = parity.build_encoder("vit_base_patch16_224", sd)
    x = synth.make_clips(4, seed=4)
    mask = synth.tube_mask(4, 0.9, seed=4)
    assert (mask.numpy() == g["mask"]).all()
    y = model(x.to(DEV), mask.to(DEV))
    assert y.shape == (4, 160, 768)
    ref = torch.from_numpy(g["tokens"].astype("float32"))
    cos = parity.row_cosine_min(y.cpu(), ref)
    rel = float((y.cpu() - ref).norm() / ref.norm())
    assert cos >= parity.TOL_COS, f"config4: worst token cosine {cos:.6f}"
    assert rel <= parity.TOL_HIDDEN_REL_L2, f"config4: rel-L2 {rel:.3e}"
    # property at BASELINE size (B=100): the first 4 clips of a 100-clip batch equal the 4-clip batch bit-for-bit
    xb = torch.cat([x, synth.make_clips(96, seed=40)]).to(DEV)
    mb = torch.cat([mask, synth.tube_mask(96, 0.9, seed=40)]).to(DEV)
    yb = model(xb, mb)
    assert yb.shape == (100, 160, 768) and torch.isfinite(yb).all()
    cos_b = parity.row_cosine_min(yb[:4].cpu(), ref)
    assert cos_b >= parity.TOL_COS, f"config4 @B=100: worst token cosine {cos_b:.6f}"


def test_config5_batch_sweep_consistency():
    """BASELINE config 5 (correctness side of the batch sweep): ViT-B logits for the same clips at batch 1, 2, 8
    all match the reference within tolerance."""
    g = parity.golden("c5_vitb_b8")
    sd = synth.make_state_dict("vit_base_patch16_224", seed=5)
    model = parity.build_classifier("vit_base_patch16_224", sd)
    x = synth.make_clips(8, seed=5).to(DEV)
    ref = g["logits"]
    parity.check_logits(model(x), ref, "config5 B=8")
    parity.check_logits(torch.cat([model(x[i:i + 1]) for i in range(8)]), ref, "config5 B=1 x8")
    parity.check_logits(torch.cat([model(x[i:i + 2]) for i in range(0, 8, 2)]), ref, "config5 B=2 x4")
    big = torch.cat([x, synth.make_clips(56, seed=50).to(DEV)])
    out = model(big)                                      # B = 64, the bench batch
    parity.check_logits(out[:8], ref, "config5 B=64 (first 8)")


def test_modules_are_drop_in():
    """Mlp / Attention / Block / PatchEmbed / FlashAttention individually against the oracle's functions."""
    from simple_tad_b200 import modeling_finetune as mf
    from simple_tad_b200.flash_attention_class import FlashAttention
    arch, seed = "vit_small_d2", 21
    D, depth, heads = synth.ARCHS[arch]
    sd = synth.make_state_dict(arch, seed=seed)
    model = parity.build_classifier(arch, sd)
    g = torch.Generator().manual_seed(5)
    x = synth.bf16_round(torch.randn(2, 392, D, generator=g))
    blk = model.blocks[0]
    p = "blocks.0."

    def rel(a, b):
        return float((a.float().cpu() - b).norm() / b.norm())

    assert rel(blk.mlp(x.to(DEV)), vit_oracle.mlp(sd, p, x)) < 1e-2
    assert rel(blk.attn(x.to(DEV)), vit_oracle.attention(sd, p, x, heads)) < 1e-2
    assert rel(blk(x.to(DEV)), vit_oracle.block(sd, 0, x, heads)) < 1e-2
    clips = synth.make_clips(1, seed=seed)
    assert rel(model.patch_embed(clips.to(DEV)), vit_oracle.patch_embed(sd, clips)) < 1e-2
    # FlashAttention.forward contract (fac:26-51): qkv [B,S,3,H,D] -> (out [B,S,H,D], None)
    qkv = synth.bf16_round(torch.randn(2, 160, 3, heads, 64, generator=g))
    out, none = FlashAttention()(qkv.to(DEV).to(torch.bfloat16))
    assert none is None and out.shape == (2, 160, heads, 64)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3)
    assert rel(out, ref) < 1e-2


def test_errors_follow_reference_conventions():
    from simple_tad_b200 import _lib
    sd = synth.make_state_dict("vit_small_d2", seed=11)
    model = parity.build_classifier("vit_small_d2", sd)
    with pytest.raises(AssertionError):      # mf:188 asserts on the image size
        model(torch.zeros(1, 3, 16, 112, 112, device=DEV))
    with pytest.raises(NotImplementedError):
        model.train()(torch.zeros(1, 3, 16, 224, 224, device=DEV))
    model.eval()
    with pytest.raises(ValueError):          # K not a multiple of 64 -> STAD_E_SHAPE
        _lib.gemm_bias_residual(torch.zeros(128, 72, device=DEV, dtype=torch.bfloat16),
                                torch.zeros(64, 72, device=DEV, dtype=torch.bfloat16))
    assert "K=72" in _lib.last_error()
