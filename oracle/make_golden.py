"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on the synthetic
weights / inputs of oracle/synth.py.  Run in the build container (the reference is not present on the GPU box):

    python -m oracle.make_golden            # all fixtures (~10 min of CPU)
    python -m oracle.make_golden c1 peaky   # a subset

Shims: `timm` is not installed, so the four symbols the reference imports from it are provided
(timm.models.layers.{drop_path,to_2tuple,trunc_normal_}, timm.models.registry.register_model); `natsort` is stubbed
for run_inference_simple.py.  Neither touches the forward math.  The script also reports the max abs difference between
the reference outputs and oracle/vit_oracle.py on the same inputs.
"""
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("STAD_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import synth, vit_oracle  # noqa: E402

HID_TOK = [0, 1, 13, 195, 196, 783, 1000, 1567]   # sampled tokens of the 1568
PIX_TOK_STEP, PIX_VAL_STEP = 11, 6               # pretrain fixtures keep pixels[:, ::11, ::6]
HID_CH = [0, 1, 2, 63, 64, 191, 255, 383]         # sampled channels (valid for D >= 384)


def install_shims(ref_path=None):
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        registry = types.ModuleType("timm.models.registry")

        def drop_path(x, drop_prob=0.0, training=False):
            if drop_prob == 0.0 or not training:
                return x
            raise RuntimeError("drop_path shim is eval-only")

        def to_2tuple(v):
            return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        _reg = {}

        def register_model(fn):
            _reg[fn.__name__] = fn
            return fn

        layers.drop_path, layers.to_2tuple, layers.trunc_normal_ = drop_path, to_2tuple, trunc_normal_
        registry.register_model, registry._model_entrypoints = register_model, _reg
        timm.models, models.layers, models.registry = models, layers, registry
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                            "timm.models.registry": registry})
    if "natsort" not in sys.modules:
        ns = types.ModuleType("natsort")
        ns.natsorted = sorted
        sys.modules["natsort"] = ns
    ref_path = ref_path or REF
    if ref_path not in sys.path:
        sys.path.insert(0, ref_path)


def ref_classifier(arch, sd, init_values=0.0):
    """The reference VisionTransformer (modeling_finetune.py) in fp32 with the naive attention path."""
    import modeling_finetune as mf
    from functools import partial
    D, depth, heads = synth.ARCHS[arch]
    if arch in mf.__dict__ and init_values == 0.0:
        model = mf.__dict__[arch](num_classes=2, all_frames=16, tubelet_size=2, use_flash_attn=False, init_scale=1.0,
                                  final_reduction="fc_norm")
    else:  # reduced-depth variant: same ctor the factories call (mf:340-342), only `depth` differs
        model = mf.VisionTransformer(patch_size=16, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                     norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16,
                                     tubelet_size=2, use_flash_attn=False, init_scale=1.0, final_reduction="fc_norm",
                                     init_values=init_values)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()


def ref_encoder(arch, sd):
    import modeling_pretrain as mp
    from functools import partial
    D, depth, heads = synth.ARCHS[arch]
    enc = mp.PretrainVisionTransformerEncoder(
        img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4,
        qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0., tubelet_size=2,
        use_flash_attn=False)   # same kwargs PretrainVisionTransformer passes down (mp:215-232)
    enc.load_state_dict(sd, strict=True)
    return enc.eval()


@torch.no_grad()
def ref_hidden(model, x):
    """Residual stream after patch-embed+pos and after each block, via forward hooks on the reference modules."""
    hs = []
    hooks = [blk.register_forward_hook(lambda m, i, o: hs.append(o.detach())) for blk in model.blocks]
    pre = model.blocks[0].register_forward_pre_hook(lambda m, i: hs.append(i[0].detach()))
    logits = model(x)
    for h in hooks + [pre]:
        h.remove()
    return logits, hs


def sample_hidden(hs):
    tok = torch.tensor(HID_TOK)
    ch = torch.tensor(HID_CH)
    samp = torch.stack([h[:, tok][:, :, ch] for h in hs])            # [L+1, B, 8, 8]
    rms = torch.stack([h.pow(2).mean((1, 2)).sqrt() for h in hs])    # [L+1, B]
    return samp.numpy(), rms.numpy()


def save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def gen_classifier(name, arch, B=None, video_T=None, n_videos=1, seed=0, peaky=1.0, use_ris=False, check_oracle=True,
                   trained_like=False, extra_clips_seed=None):
    D, depth, heads = synth.ARCHS[arch]
    sd = (synth.make_trained_like_state_dict(arch, seed=seed) if trained_like else
          synth.make_state_dict(arch, seed=seed, peaky=peaky))
    if video_T is None and extra_clips_seed is not None:
        x = torch.cat([synth.make_clips(8, seed=seed), synth.make_clips(B - 8, seed=extra_clips_seed)])
    elif video_T is None:
        x = synth.make_clips(B, seed=seed)
    else:
        x = torch.cat([synth.windows_from_video(synth.make_video(video_T, seed=seed + v)) for v in range(n_videos)])
    model = ref_classifier(arch, sd, init_values=0.1 if trained_like else 0.0)
    t0 = time.time()
    logits, hs = [], None
    for i in range(0, x.shape[0], 4):
        if i == 0:
            lg, hs = ref_hidden(model, x[:4])
        else:
            with torch.no_grad():
                lg = model(x[i:i + 4])
        logits.append(lg)
    logits = torch.cat(logits)
    print(f"{name}: reference {arch} on {tuple(x.shape)} in {time.time() - t0:.1f}s; logits[0]={logits[0].tolist()}")
    samp, rms = sample_hidden(hs)
    extra = {}
    if use_ris:
        import run_inference_simple as ris
        m2 = ris.get_video_vit_small(with_flash=False) if D == 384 else ris.get_video_vit_base(with_flash=False)
        m2.load_state_dict(sd, strict=True)
        with torch.no_grad():
            probs = m2.eval()(x)
        assert torch.allclose(probs, logits.softmax(-1), atol=1e-6), "run_inference_simple model disagrees with modeling_finetune"
        extra["probs_ris"] = probs.numpy()
    if check_oracle:
        o = vit_oracle.vit_forward(sd, x[:4], heads)
        print(f"  oracle vs reference: max|dlogit| = {(o - logits[:4]).abs().max():.3e}")
    save(name, logits=logits.numpy(), probs=logits.softmax(-1).numpy(), hidden_samples=samp, hidden_rms=rms,
         hid_tok=np.array(HID_TOK), hid_ch=np.array(HID_CH), meta=np.array([seed, x.shape[0], peaky]), **extra)


def gen_encoder(name, arch, B, ratio=0.9, seed=0):
    D, depth, heads = synth.ARCHS[arch]
    sd = synth.make_state_dict(arch, seed=seed, encoder=True)
    x = synth.make_clips(B, seed=seed)
    mask = synth.tube_mask(B, ratio, seed=seed)
    enc = ref_encoder(arch, sd)
    t0 = time.time()
    with torch.no_grad():
        y = enc(x, mask)
    print(f"{name}: reference encoder {arch} {tuple(x.shape)} mask {ratio} -> {tuple(y.shape)} in {time.time() - t0:.1f}s")
    o = vit_oracle.encoder_forward(sd, x, mask, heads)
    print(f"  oracle vs reference: max|d| = {(o - y).abs().max():.3e}")
    save(name, tokens=y.numpy().astype(np.float16), token_norm=y.norm(dim=-1).numpy(),
         mask=mask.numpy(), meta=np.array([seed, B, ratio]))


def gen_pretrain(name, arch, B, ratio=0.9, seed=0, decoder_depth=4):
    """Full PretrainVisionTransformer (encoder + decoder, mp:183-291) of the unmodified reference."""
    import modeling_pretrain as mp
    from functools import partial
    D, depth, heads = synth.ARCHS[arch]
    Dd, dheads = synth.DECODERS[arch]
    sd = synth.make_pretrain_state_dict(arch, seed=seed, decoder_depth=decoder_depth)
    model = mp.PretrainVisionTransformer(
        img_size=224, patch_size=16, encoder_embed_dim=D, encoder_depth=depth, encoder_num_heads=heads,
        encoder_num_classes=0, decoder_num_classes=1536, decoder_embed_dim=Dd, decoder_num_heads=dheads,
        decoder_depth=decoder_depth, mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
        use_flash_attn=False)   # the kwargs of the factories mp:293-363 (+ decoder_depth, run_mae_pretraining.py)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model.eval()
    x = synth.make_clips(B, seed=seed)
    mask = synth.tube_mask(B, ratio, seed=seed)
    t0 = time.time()
    with torch.no_grad():
        y = model(x, mask)
    print(f"{name}: reference pretrain {arch} {tuple(x.shape)} mask {ratio} -> {tuple(y.shape)} in {time.time() - t0:.1f}s")
    o = vit_oracle.pretrain_forward(sd, x, mask, heads, dheads)
    print(f"  oracle vs reference: max|d| = {(o - y).abs().max():.3e} (|y| max {y.abs().max():.3f})")
    # sampled tokens x sampled pixel values in fp32 + the norm of EVERY predicted token (keeps the fixture small)
    save(name, pixels_sample=y[:, ::PIX_TOK_STEP, ::PIX_VAL_STEP].numpy(), pixel_norm=y.norm(dim=-1).numpy(),
         pixel_mean=y.mean(dim=1).numpy(), mask=mask.numpy(), meta=np.array([seed, B, ratio, decoder_depth]))


def gen_masks(name):
    """TubeMaskingGenerator of the unmodified reference (masking_generator.py) under fixed np.random seeds."""
    import masking_generator as mg
    out = {}
    for seed, ratio in ((0, 0.9), (1, 0.75), (2, 0.5)):
        np.random.seed(seed)
        gen = mg.TubeMaskingGenerator((8, 14, 14), ratio)
        out[f"mask_s{seed}"] = np.stack([gen() for _ in range(3)])
        out[f"ratio_s{seed}"] = np.array(ratio)
        np.random.seed(seed)
        mine = np.stack([vit_oracle.TubeMaskingGenerator((8, 14, 14), ratio)() for _ in range(3)])
        assert np.array_equal(mine, out[f"mask_s{seed}"]), "oracle TubeMaskingGenerator differs from the reference"
    print(f"{name}: reference TubeMaskingGenerator, 3 seeds x 3 draws; oracle restatement identical")
    save(name, **out)


def gen_eval_metrics(name):
    """anaysis/metrics.py `calculate_MORE_metrics` of the unmodified reference (scikit-learn) on synthetic scores."""
    from anaysis import metrics as ref_metrics
    from oracle import metrics_oracle as mo
    out = {}
    for seed, n in ((0, 5000), (1, 257)):
        logits, labels = mo.synthetic_scores(n, seed=seed)
        probs = torch.from_numpy(logits).softmax(-1).numpy()     # fp32, as final_test holds them (eff:461-462)
        r = ref_metrics.calculate_MORE_metrics(probs[:, 1], labels)
        (acc, prec, rec, f1, ap, auroc, confmat, _pr, _roc, mcc_t, p_t, r_t, acc_t, f1_t) = r
        mine, _ = mo.thresholded(probs[:, 1], labels)
        for k, ref_list in (("mcc", mcc_t), ("precision", p_t), ("recall", r_t), ("acc", acc_t), ("f1", f1_t)):
            assert np.allclose(mine[k], ref_list, atol=1e-12), f"oracle thresholded {k} differs from the reference"
        b = mo.torchmetrics_binned(probs[:, 1], labels)
        print(f"{name}[seed {seed}, n {n}]: sklearn auroc {auroc:.6f} ap {ap:.6f}; binned(101) auroc {b['auroc']:.6f} ap {b['ap']:.6f}")
        # the group summaries of anaysis/metrics.py:19-125 (exact AP / AUROC over every score; false-negative shares of
        # the positives)
        pos = probs[labels == 1, 1]
        ones = np.ones(len(pos), dtype=np.int64)
        out.update({f"calc_metrics_s{seed}": np.array(ref_metrics.calculate_metrics(probs[:, 1], labels)),
                    f"fn_group_s{seed}": np.array([ref_metrics.calculate_fn_group(pos, ones)]),
                    f"fn_group_thr_s{seed}": np.array(ref_metrics.calculate_fn_group_thresholds(pos, ones))})
        out.update({f"probs_s{seed}": probs, f"labels_s{seed}": labels, f"mcc_s{seed}": np.array(mcc_t),
                    f"precision_s{seed}": np.array(p_t), f"recall_s{seed}": np.array(r_t), f"acc_s{seed}": np.array(acc_t),
                    f"f1_s{seed}": np.array(f1_t), f"at05_s{seed}": np.array([acc, prec, rec, f1]),
                    f"confmat_s{seed}": np.array(confmat), f"sk_auroc_ap_s{seed}": np.array([auroc, ap])})
    save(name, **out)


def gen_resize(name):
    """cv2.resize(..., INTER_CUBIC) as the reference's callers invoke it (ri:79-80, dota.py:347-348), OpenCV's own C++
    path (IPP off), on synthetic uint8 frames; also records how far the default IPP dispatch is from that path."""
    import cv2
    from oracle import resize_oracle as ro
    out = {"cv2_version": np.array(cv2.__version__)}
    for i, (h, w, dh, dw) in enumerate(((720, 1280, 224, 224), (360, 640, 224, 224), (100, 60, 37, 23))):
        img = ro.synthetic_frame(h, w, seed=i)
        ipp = cv2.ipp.useIPP()
        with_ipp = cv2.resize(img, dsize=(dw, dh), interpolation=cv2.INTER_CUBIC)
        cv2.ipp.setUseIPP(False)
        ref = cv2.resize(img, dsize=(dw, dh), interpolation=cv2.INTER_CUBIC)
        cv2.ipp.setUseIPP(ipp)
        mine = ro.resize_cubic_u8(img, dh, dw)
        d = np.abs(ref.astype(int) - mine.astype(int))
        d_ipp = np.abs(ref.astype(int) - with_ipp.astype(int))
        print(f"{name}[{h}x{w} -> {dh}x{dw}]: oracle vs cv2 (IPP off) mismatch {(d > 0).mean():.2e} max {d.max()}; "
              f"cv2 IPP on vs off mismatch {(d_ipp > 0).mean():.2e} max {d_ipp.max()}")
        assert d.max() <= 1 and (d > 0).mean() < 1e-3
        out.update({f"shape_{i}": np.array([h, w, dh, dw]), f"cv2_{i}": ref,
                    f"rates_{i}": np.array([(d > 0).mean(), (d_ipp > 0).mean()])})
    save(name, **out)


# ---- the other forms of the classifier: final_reduction 'cls' / 'none' and the MVD / UMT siblings -------------------------
# name: (family, arch, B, seed, kwargs of the reference ctor beyond the common ones, clip frames, image size)
VARIANTS = {
    "var_mf_cls_vits_d2_b2": ("mf", "vit_small_d2", 2, 21, dict(final_reduction="cls"), 16, 224),
    "var_mf_none_vits_d2_b2": ("mf", "vit_small_d2", 2, 22, dict(final_reduction="none"), 16, 224),
    "var_mvd_vits_d2_b2": ("mvd", "vit_small_d2", 2, 23, dict(), 16, 224),
    "var_mvd_clstok_vits_d2_b2": ("mvd", "vit_small_d2", 2, 24, dict(use_cls_token=True), 16, 224),
    "var_mvd_clstok_clsred_vits_d2_b2": ("mvd", "vit_small_d2", 2, 25, dict(use_cls_token=True, final_reduction="cls"), 16, 224),
    "var_umt_t1f8_vitb_d2_b2": ("umt", "vit_base_d2", 2, 26, dict(tubelet_size=1, all_frames=8), 8, 224),
    "var_umt_t1f16_vits_d2_b1": ("umt", "vit_small_d2", 1, 27, dict(tubelet_size=1, all_frames=16), 16, 224),
    "var_umt_384_vits_d2_b1": ("umt", "vit_small_d2", 1, 28, dict(tubelet_size=1, all_frames=8, img_size=384), 8, 384),
    "var_mvd_vitb_b4": ("mvd", "vit_base_patch16_224", 4, 29, dict(), 16, 224),
}


def _ref_module(family):
    """The unmodified reference module of a family, imported from its file under a private name (the three files share
    the name modeling_finetune.py)."""
    import importlib.util
    rel = {"mf": "modeling_finetune.py", "mvd": "other_models/MVD/modeling_finetune.py",
           "umt": "other_models/UMT/modeling_finetune.py"}[family]
    key = "_stad_ref_" + family
    if key not in sys.modules:
        spec = importlib.util.spec_from_file_location(key, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[key] = mod
        spec.loader.exec_module(mod)
    return sys.modules[key]


def variant_setup(name):
    """(reference-ctor kwargs, state dict, clips, oracle position table) of a VARIANTS entry — shared with the tests."""
    family, arch, B, seed, extra, frames, img = VARIANTS[name]
    D, depth, heads = synth.ARCHS[arch]
    tubelet = extra.get("tubelet_size", 2)
    red = extra.get("final_reduction", "fc_norm")
    sd = synth.make_variant_state_dict(arch, seed=seed, tubelet=tubelet, final_reduction=red,
                                       cls_token=extra.get("use_cls_token", False))
    x = synth.make_clips(B, seed=seed, frames=frames, img=img)
    n_tok = (frames // tubelet) * (img // 16) ** 2
    if family == "mvd":
        pos = vit_oracle.sincos_3d_table(D, img // 16, frames // tubelet)
    elif family == "umt":
        pos = vit_oracle.umt_table(n_tok, D, frames // tubelet)
    else:
        pos = vit_oracle.sinusoid_table(n_tok, D)
    return family, arch, extra, sd, x, pos, red


def gen_variant(name):
    from functools import partial
    family, arch, extra, sd, x, pos, red = variant_setup(name)
    D, depth, heads = synth.ARCHS[arch]
    mod = _ref_module(family)
    kw = dict(patch_size=16, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16, tubelet_size=2,
              use_flash_attn=False, init_scale=1.0, final_reduction="fc_norm")
    kw.update(extra)
    model = mod.VisionTransformer(**kw)
    sd_load = dict(sd)
    if isinstance(model.pos_embed, torch.nn.Parameter):   # UMT with an interpolated table: the table is a parameter
        sd_load["pos_embed"] = model.pos_embed.detach().clone()
    res = model.load_state_dict(sd_load, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model.eval()
    t0 = time.time()
    with torch.no_grad():
        feat = model.forward_features(x)
        logits = model(x)
    ref_pos = model.pos_embed.detach()
    print(f"{name}: reference {family} {arch} {extra} on {tuple(x.shape)} -> logits {tuple(logits.shape)} in "
          f"{time.time() - t0:.1f}s; logits.flat[:2]={logits.flatten()[:2].tolist()}")
    print(f"  position table {tuple(ref_pos.shape)}: oracle vs reference max|d| = {(pos - ref_pos).abs().max():.3e}")
    o_logits, o_feat = vit_oracle.vit_forward_variant(sd, x, heads, pos, final_reduction=red, mvd=family == "mvd")
    print(f"  oracle vs reference: max|dlogit| = {(o_logits - logits).abs().max():.3e}, max|dfeat| = "
          f"{(o_feat - feat).abs().max():.3e}")
    n_tok = ref_pos.shape[1]
    tok = torch.tensor([t for t in HID_TOK if t < n_tok] + [n_tok - 1])
    ch = torch.tensor(HID_CH)
    feat_s = feat[..., ch] if feat.dim() == 2 else feat[:, ::97][..., ch]
    save(name, logits=logits.numpy(), probs=logits.softmax(-1).numpy(), feat_samples=feat_s.numpy(),
         feat_norm=feat.norm(dim=-1).numpy(), pos_samples=ref_pos[0][tok][:, ch].numpy(), pos_tok=tok.numpy(),
         pos_sum=np.array([ref_pos.double().sum().item(), ref_pos.double().abs().sum().item()]),
         hid_ch=np.array(HID_CH))


SEQ_CASES = [  # (timesteps_nb, input_frequency, seq_frequency, seq_length, step)
    (100, 10, 10, 16, 1), (16, 10, 10, 16, 1), (15, 10, 10, 16, 1), (100, 10, 10, 16, 10), (103, 10, 10, 16, 7),
    (150, 30, 10, 16, 1), (150, 30, 10, 16, 3), (151, 30, 10, 16, 10), (46, 30, 10, 16, 1), (45, 30, 10, 16, 1),
    (90, 30, 10, 8, 10), (64, 20, 10, 8, 5), (200, 30, 30, 16, 4), (107, 10, 10, 16, 10), (159, 30, 10, 16, 20),
]


def gen_sequences(name):
    """RegularSequencer.get_sequences of the unmodified dataset/sequencing.py (dota.py:209-213, dada.py:173-177) on a
    grid of video lengths / frame rates / steps: pins simple_tad_b200.sequencing.window_plan."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_stad_ref_sequencing", os.path.join(REF, "dataset", "sequencing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {"cases": np.array(SEQ_CASES)}
    for i, (T, fin, fseq, length, step) in enumerate(SEQ_CASES):
        seqs = mod.RegularSequencer(seq_frequency=fseq, seq_length=length, step=step).get_sequences(T, fin)
        out[f"seq_{i}"] = np.zeros((0, length), dtype=np.int64) if seqs is None else np.array(seqs, dtype=np.int64)
        ws = mod.RegularSequencerWithStart(seq_frequency=fseq, seq_length=length, step=step).get_sequences(T, fin)
        out[f"seqws_{i}"] = np.zeros((0, length), dtype=np.int64) if ws is None else np.array(ws, dtype=np.int64)
        print(f"  {SEQ_CASES[i]}: {0 if seqs is None else len(seqs)} windows (+{0 if ws is None else len(ws) - len(seqs)} with start)")
    save(name, **out)


def main():
    install_shims()
    torch.set_num_threads(os.cpu_count())
    want = set(sys.argv[1:])

    def on(k):
        return not want or k in want
    # fast fixtures (also exercised by the CPU test-suite)
    if on("small"):
        gen_classifier("small_vits_d2_b2", "vit_small_d2", B=2, seed=11)
        gen_encoder("small_enc_vitb_d2_b2", "vit_base_d2", B=2, seed=12)
        gen_pretrain("small_mae_vits_d2_b2", "vit_small_d2", B=2, seed=14, decoder_depth=2)
        gen_masks("tube_masks")
        gen_eval_metrics("eval_metrics")
        gen_resize("resize_cubic")
    if want and "eval_metrics" in want:
        gen_eval_metrics("eval_metrics")
    if on("peaky"):
        gen_classifier("peaky_vits_d2_b2", "vit_small_d2", B=2, seed=13, peaky=3.0)
    # the five BASELINE.json configs
    if on("c1"):
        gen_classifier("c1_vits_b4", "vit_small_patch16_224", B=4, seed=1, use_ris=True)
    if on("c2"):
        gen_classifier("c2_vitb_video100", "vit_base_patch16_224", video_T=100, seed=2)
    if on("c3"):
        gen_classifier("c3_vitl_2x20", "vit_large_patch16_224", video_T=20, n_videos=2, seed=3)
    if on("c4"):
        gen_encoder("c4_enc_vitb_b4", "vit_base_patch16_224", B=4, seed=4)
    if on("mae"):
        gen_pretrain("c4_mae_vitb_b2", "vit_base_patch16_224", B=2, seed=6, decoder_depth=4)
    if on("c5"):
        gen_classifier("c5_vitb_b8", "vit_base_patch16_224", B=8, seed=5)
    if on("c5big"):
        # the bench batch: all 64 clips of a B = 64 forward against the reference (the first 8 are c5_vitb_b8's clips)
        gen_classifier("c5_vitb_b64", "vit_base_patch16_224", B=64, seed=5, check_oracle=False, extra_clips_seed=50)
    if on("c3big"):
        # config 3 at a size that exercises several batches: ViT-L, 2 videos x 47 frames = 2 x 32 windows
        gen_classifier("c3_vitl_2x47", "vit_large_patch16_224", video_T=47, n_videos=2, seed=3, check_oracle=False)
    if on("trained"):
        # trained-like statistics (outlier channels, shifted rows, wide LayerNorm gamma, layer scale): see
        # synth.make_trained_like_state_dict
        gen_classifier("trained_vits_d2_b2", "vit_small_d2", B=2, seed=31, trained_like=True)
        gen_classifier("trained_vitb_b4", "vit_base_patch16_224", B=4, seed=32, trained_like=True)
    if on("sequencer"):
        gen_sequences("sequencer")
    for name in VARIANTS:
        if on("variants") or (want and name in want):
            gen_variant(name)


if __name__ == "__main__":
    main()
