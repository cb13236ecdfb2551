"""CPU: host-side logic of the drop-in modules — registry, checkpoint naming contract, weight preparation
(LayerNorm folding), position table, token index lists.  No kernel is launched."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth, vit_oracle
from simple_tad_b200 import modeling_finetune as mf
from simple_tad_b200 import modeling_pretrain as mp
from simple_tad_b200.registry import create_model, is_model, list_models


def test_registry_has_every_reference_factory():
    for name in ("vit_small_patch16_224", "vit_base_patch16_224", "vit_base_patch16_384", "vit_large_patch16_224",
                 "vit_large_patch16_384", "vit_large_patch16_512", "vit_huge_patch16_224",
                 "pretrain_videomae_small_patch16_224", "pretrain_videomae_base_patch16_224",
                 "pretrain_videomae_large_patch16_224", "pretrain_videomae_huge_patch16_224"):
        assert is_model(name), name
    assert len(list_models()) >= 11
    with pytest.raises(RuntimeError):
        create_model("vit_tiny_does_not_exist")


def test_create_model_call_site_of_run_frame_finetuning():
    """rff:374-389 passes drop_block_rate=None, which the ctor does not accept: create_model must drop None kwargs."""
    m = create_model("vit_small_patch16_224", pretrained=False, num_classes=2, all_frames=16, tubelet_size=2,
                     fc_drop_rate=0.0, drop_rate=0.0, drop_path_rate=0.1, attn_drop_rate=0.0, drop_block_rate=None,
                     use_checkpoint=False, final_reduction="fc_norm", init_scale=0.001, use_flash_attn=True)
    assert sum(p.numel() for p in m.parameters()) == 21_880_706   # ViT-S 21.88 M params
    assert m.get_num_layers() == 12 and m.num_heads == 6
    assert m.patch_embed.patch_size == (16, 16) and m.patch_embed.num_patches == 1568 and m.patch_embed.tubelet_size == 2
    assert m.pos_embed.shape == (1, 1568, 384) and "pos_embed" not in m.state_dict()
    assert m.no_weight_decay() == {"pos_embed", "cls_token"}
    assert m.default_cfg["num_classes"] == 400
    assert float(m.head.weight.abs().max()) < 1e-3  # init_scale applied (mf:281-283)
    assert isinstance(m.blocks[3].drop_path, mf.DropPath) and isinstance(m.blocks[0].drop_path, torch.nn.Identity)


def test_state_dict_contract_matches_reference_names_and_shapes():
    for arch, factory in (("vit_small_patch16_224", mf.vit_small_patch16_224), ("vit_base_patch16_224", mf.vit_base_patch16_224)):
        sd = synth.make_state_dict(arch)
        m = factory(num_classes=2)
        own = m.state_dict()
        assert set(own) == set(sd)
        for k in sd:
            assert own[k].shape == sd[k].shape, k
        assert "blocks.0.attn.qkv.bias" not in own and "blocks.0.attn.q_bias" in own  # K has no bias (mf:69-75)
    full = mp.pretrain_videomae_base_patch16_224(decoder_depth=4)
    sde = synth.make_state_dict("vit_base_patch16_224", encoder=True)
    assert set(full.encoder.state_dict()) == set(sde) and "norm.weight" in sde and "head.weight" not in sde
    # full pre-training model (mp:183-258): encoder.*, decoder.*, encoder_to_decoder.weight (no bias), mask_token
    sdp = synth.make_pretrain_state_dict("vit_base_patch16_224", decoder_depth=4)
    own = full.state_dict()
    assert set(own) == set(sdp)
    for k in sdp:
        assert own[k].shape == sdp[k].shape, k
    assert "pos_embed" not in own and full.pos_embed.shape == (1, 1568, 384)
    assert own["decoder.head.weight"].shape == (1536, 384) and own["mask_token"].shape == (1, 1, 384)
    assert sum(p.numel() for p in full.parameters()) == 94_210_944  # VideoMAE-B pre-training model, decoder depth 4
    assert full.no_weight_decay() == {"pos_embed", "cls_token", "mask_token"}
    small = mp.pretrain_videomae_small_patch16_224(decoder_depth=4)
    assert small.decoder.embed_dim == 192 and small.decoder.num_heads == 3


def test_sinusoid_table_is_bit_identical_to_the_oracle():
    a = mf.get_sinusoid_encoding_table(1568, 384)
    b = vit_oracle.sinusoid_table(1568, 384)
    assert a.dtype == torch.float32 and a.shape == (1, 1568, 384)
    assert torch.equal(a, b)


def test_layernorm_fold_is_algebraically_exact():
    """LN(x) W^T + b == rstd * (x W'^T - mean * colsum(W')) + b'   with W' = W diag(gamma), b' = b + W beta."""
    torch.manual_seed(0)
    D = 384
    blk = mf.Block(D, 6, qkv_bias=True, init_values=0., norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6)).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn_like(p) * 0.05 + (1.0 if p.dim() == 1 and p is blk.norm1.weight else 0.0))
    pk = blk.packed("cpu")
    x = torch.randn(50, D) * 2 + 0.5
    mean = x.mean(1, keepdim=True)
    rstd = (x.var(1, unbiased=False, keepdim=True) + 1e-6).rsqrt()
    ref = F.linear(F.layer_norm(x, (D,), blk.norm1.weight, blk.norm1.bias, 1e-6), blk.attn.qkv.weight, blk.attn.packed_qkv_bias())
    w = pk["w_qkv"].float()
    got = rstd * (x @ w.t() - mean * pk["cs_qkv"]) + pk["b_qkv"]
    # only the bf16 rounding of W' separates the two
    assert float((got - ref).abs().max()) < 2e-2 * float(ref.abs().max())
    assert torch.equal(pk["cs_qkv"], w.sum(1))
    D3 = 3 * D
    assert pk["w_qkv"].shape == (D3, D) and pk["w_fc1"].shape == (4 * D, D) and pk["w_fc2"].shape == (D, 4 * D)
    # K rows get no bias except the beta term
    k_bias = pk["b_qkv"][D:2 * D]
    assert torch.allclose(k_bias, blk.attn.qkv.weight[D:2 * D] @ blk.norm1.bias, atol=1e-5)


def test_layer_scale_is_folded_into_proj_and_fc2():
    blk = mf.Block(384, 6, qkv_bias=True, init_values=0.1, norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6)).eval()
    pk = blk.packed("cpu")
    assert torch.allclose(pk["w_proj"].float(), (blk.attn.proj.weight * 0.1).to(torch.bfloat16).float())
    assert torch.allclose(pk["b_fc2"], blk.mlp.fc2.bias * 0.1)


def test_visible_token_indices_match_boolean_indexing_order():
    mask = synth.tube_mask(5, 0.9, seed=3)
    idx, n = mp.visible_token_indices(mask)
    assert n == 160 and idx.dtype == torch.int32
    assert torch.equal(idx, vit_oracle.visible_indices(mask))
    idx75, n75 = mp.visible_token_indices(synth.tube_mask(2, 0.75, seed=1))
    assert n75 == 392


def test_masked_token_indices_complement_the_visible_ones():
    mask = synth.tube_mask(3, 0.9, seed=5)
    vis, nv = mp.visible_token_indices(mask)
    msk, nm = mp.masked_token_indices(mask)
    assert nv == 160 and nm == 1408 and msk.dtype == torch.int32
    for b in range(3):
        assert torch.equal(msk[b].long(), mask[b].nonzero().flatten())          # row-major order of x[mask] (mp:286)
        assert sorted(vis[b].tolist() + msk[b].tolist()) == list(range(1568))


def test_norm_fold_of_the_pixel_head_is_exact():
    """decoder.norm folded into decoder.head (and encoder.norm into encoder_to_decoder, which has no bias)."""
    torch.manual_seed(1)
    norm = torch.nn.LayerNorm(192, eps=1e-6)
    lin = torch.nn.Linear(192, 1536)
    e2d = torch.nn.Linear(192, 128, bias=False)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(192))
        norm.bias.copy_(0.05 * torch.randn(192))
    x = torch.randn(40, 192) * 1.5 + 0.3
    mean = x.mean(1, keepdim=True)
    rstd = (x.var(1, unbiased=False, keepdim=True) + 1e-6).rsqrt()
    for layer in (lin, e2d):
        w, b, cs = mp._fold_norm_linear(norm, layer.weight, layer.bias, "cpu")
        ref = layer(norm(x))
        got = rstd * (x @ w.float().t() - mean * cs) + b
        assert float((got - ref).abs().max()) < 2e-2 * float(ref.abs().max())


def test_metrics_host_math_from_histogram_matches_the_oracle():
    """simple_tad_b200.metrics: suffix sums of the threshold histogram -> confusion counts -> every metric, against the
    oracle's per-threshold loops (the histogram the kernel produces is emulated here with numpy.searchsorted)."""
    from oracle import metrics_oracle as mo
    from simple_tad_b200 import metrics as M
    assert M.THRESHOLDS == mo.THRESHOLDS and len(M.THRESHOLDS) == 101
    logits, labels = mo.synthetic_scores(3000, seed=5)
    probs = torch.from_numpy(logits).softmax(-1).numpy()
    thr32 = np.asarray(M.THRESHOLDS, dtype=np.float64).astype(np.float32)
    bins = np.searchsorted(thr32, probs[:, 1], side="right")           # number of thresholds <= p
    hist = np.zeros((2, 102), dtype=np.int64)
    np.add.at(hist, (labels, bins), 1)
    c = M.counts_from_hist(hist)
    th, counts = mo.thresholded(probs[:, 1], labels, thr32)
    for i, k in enumerate(("tn", "fp", "fn", "tp")):
        assert np.array_equal(c[k], counts[:, i]), k
    got = M.thresholded_metrics(c)
    for k in ("mcc", "precision", "recall", "acc", "f1"):
        np.testing.assert_allclose(got[k], th[k], atol=1e-12, rtol=0)
    b, ob = M.binned_curves(c), mo.torchmetrics_binned(probs[:, 1], labels)
    assert abs(b["auroc"] - ob["auroc"]) < 1e-12 and abs(b["ap"] - ob["ap"]) < 1e-12
    np.testing.assert_allclose(b["roc_curve"][0], ob["fpr"], atol=1e-15)
    np.testing.assert_allclose(b["pr_curve"][0], ob["precision"], atol=1e-15)
    pred = probs[:, 1] > probs[:, 0]
    conf = [int((~pred & (labels == 0)).sum()), int((pred & (labels == 0)).sum()), int((~pred & (labels == 1)).sum()),
            int((pred & (labels == 1)).sum())]
    a, oa = M.argmax_metrics(conf), mo.argmax_metrics(probs, labels)
    assert a == oa
    lines = M.stats_lines({**a, **b})
    assert lines[1].startswith("mAP: ") and "auroc: " in lines[1] and lines[3].startswith("Confmat: ")


def test_inference_only_and_cuda_only():
    m = mf.vit_small_patch16_224(num_classes=2)
    with pytest.raises((NotImplementedError, RuntimeError)):
        m.train()(torch.zeros(1, 3, 16, 224, 224))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(torch.zeros(1, 3, 16, 224, 224))
    with pytest.raises(RuntimeError, match="CUDA"):  # 'cls' / 'none' are supported; still no CPU path
        mf.VisionTransformer(embed_dim=384, depth=1, num_heads=6, num_classes=2, final_reduction="cls").eval().prepare("cpu")
    with pytest.raises(NotImplementedError, match="head_dim"):
        mf.vit_huge_patch16_224(num_classes=2).eval().prepare("cpu")


def test_exact_group_metrics_match_reference_metrics_py():
    """anaysis/metrics.py:19-125 — calculate_metrics (exact AP / AUROC over every score, metrics at 0.5) and the
    false-negative shares of a positive group — against the outputs of the unmodified reference (scikit-learn) stored
    in tests/golden/eval_metrics.npz; plus the written one-class fallback (auc = -10 - label)."""
    import numpy as np
    from simple_tad_b200 import metrics as M
    from tests import parity
    g = parity.golden("eval_metrics")
    for seed in (0, 1):
        probs, labels = g[f"probs_s{seed}"], g[f"labels_s{seed}"]
        got = np.array(M.calculate_metrics(probs[:, 1], labels))
        assert np.allclose(got, g[f"calc_metrics_s{seed}"], rtol=0, atol=1e-12)
        auroc, ap = M.exact_auroc_ap(probs[:, 1], labels)
        assert abs(auroc - g[f"sk_auroc_ap_s{seed}"][0]) <= 1e-12 and abs(ap - g[f"sk_auroc_ap_s{seed}"][1]) <= 1e-12
        pos = probs[labels == 1, 1]
        ones = np.ones(len(pos), dtype=np.int64)
        assert abs(M.calculate_fn_group(pos, ones) - g[f"fn_group_s{seed}"][0]) <= 1e-15
        thr = np.array(M.calculate_fn_group_thresholds(pos, ones))
        assert np.array_equal(thr, g[f"fn_group_thr_s{seed}"])
        # the same shares from the count table of the device epilogue: fn / (fn + tp) of an all-positive group
        thr32 = np.asarray(M.THRESHOLDS, dtype=np.float64).astype(np.float32)
        hist = np.zeros((2, 102), dtype=np.int64)
        np.add.at(hist, (ones, np.searchsorted(thr32, pos, side="right")), 1)
        c = M.counts_from_hist(hist)
        assert np.allclose(c["fn"] / (c["fn"] + c["tp"]), thr, rtol=0, atol=1e-15)
        assert M.calculate_metrics(pos, ones)[5] == -11 and M.calculate_metrics(probs[labels == 0, 1], np.zeros(int((labels == 0).sum())))[5] == -10


def test_run_inference_simple_dropin_structure():
    """run_inference_simple.py:279-407: VisionTransformerInfer and its two factories keep the checkpoint contract of the
    classifier (same keys / shapes), scale the head by init_scale = 0.001, and have no CPU path; prepare_image is the
    arithmetic of ris:18-37."""
    import numpy as np
    from simple_tad_b200 import run_inference_simple as ris
    small, base = ris.get_video_vit_small(with_flash=False), ris.get_video_vit_base(with_flash=True)
    ref_small = mf.vit_small_patch16_224(num_classes=2)
    assert {k: tuple(v.shape) for k, v in small.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in ref_small.state_dict().items()}
    assert base.embed_dim == 768 and len(base.blocks) == 12 and base.num_heads == 12 and small.num_heads == 6
    assert float(small.head.weight.abs().max()) < 1e-3 and small.final_reduction == "fc_norm"
    with pytest.raises(RuntimeError, match="CUDA"):
        small.eval()(torch.zeros(1, 3, 16, 224, 224))
    g = np.random.default_rng(0)
    img = g.integers(0, 256, size=(7, 9, 3), dtype=np.uint8)           # BGR, as cv2.imread returns it
    got = ris.prepare_image(img, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    want = torch.from_numpy(img[:, :, ::-1].transpose(2, 0, 1).copy()).float().div(255.0)
    want = (want - torch.tensor((0.485, 0.456, 0.406)).view(3, 1, 1)) / torch.tensor((0.229, 0.224, 0.225)).view(3, 1, 1)
    assert got.shape == (3, 7, 9) and got.dtype == torch.float32 and torch.equal(got, want)
    with pytest.raises(TypeError):
        ris.prepare_image(img[:, :, 0], (0.5,), (0.5,))
    assert sorted(["f10.png", "f2.png", "f1.png"], key=ris._natural_key) == ["f1.png", "f2.png", "f10.png"]


def test_frame_folder_loop_feeds_the_window_like_the_reference(tmp_path, monkeypatch):
    """run_inference_simple.py:428-463 / run_inference.py:69-109: the first 16 images fill the window, the next 16 are
    skipped (`if i < 16: continue`), then one image per prediction, reported under the loop index."""
    import cv2
    import numpy as np
    from simple_tad_b200 import run_inference_simple as ris
    for k in range(40):
        cv2.imwrite(str(tmp_path / f"frame_{k}.png"), np.full((224, 224, 3), k, dtype=np.uint8))
    (tmp_path / "notes.txt").write_text("not an image")
    pushed = []

    class _Scorer:
        H, W = 224, 224

        def __init__(self, model, **kw):
            assert kw["bgr"] is True

        def push(self, frame):
            assert frame.dtype == torch.uint8 and tuple(frame.shape) == (224, 224, 3)
            pushed.append(int(frame[0, 0, 0]))
            if len(pushed) < 16:
                return None
            v = float(pushed[-1])
            return torch.tensor([0.0, v]), torch.tensor([1.0 - v / 100.0, v / 100.0])

    monkeypatch.setattr(ris, "StreamingScorer", _Scorer)
    got = list(ris.score_frame_folder(object(), str(tmp_path)))
    assert pushed == list(range(16)) + list(range(32, 40))          # natural order: frame_2 before frame_10
    assert [i for i, _ in got] == [15] + list(range(16, 24))
    assert got[0][1] == pytest.approx(0.15) and got[1][1] == pytest.approx(0.32) and got[-1][1] == pytest.approx(0.39)
