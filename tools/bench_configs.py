"""Throughput of the other BASELINE configs on one GPU (bench.py covers config 2): ViT-S / ViT-B / ViT-L classifier
forward at a given batch, the masked DAPT encoder and the full MAE pre-training forward (ViT-B, mask 0.9).  CUDA events
around `iters` forwards after warm-up; inputs rotate over 3 buffers (activations per step exceed L2).  One JSON line per
configuration.   python tools/bench_configs.py [--iters 10]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import synth_data as synth  # noqa: E402  (synthetic weights / inputs)
from simple_tad_b200 import modeling_finetune as mf, modeling_pretrain as mp  # noqa: E402
from simple_tad_b200.masking_generator import TubeMaskingGenerator, batch_masks  # noqa: E402

GF = {"vit_small_patch16_224": 113.76, "vit_base_patch16_224": 360.69, "vit_large_patch16_224": 1193.67}


def timed(fn, iters, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def siblings(iters, dev, peak):
    """The sibling classifiers at ViT-B width: MVD with / without its class token, UMT on 8 and 16 frames
    with tubelet 1, and the per-token 'none' reduction.  FLOPs per clip from SURVEY §8d's formula with the model's own
    token count S and patch-embed K."""
    from functools import partial
    from simple_tad_b200.other_models.MVD import modeling_finetune as mvd
    from simple_tad_b200.other_models.UMT import modeling_finetune as umt
    D, L = 768, 12
    common = dict(patch_size=16, embed_dim=D, depth=L, num_heads=12, mlp_ratio=4, qkv_bias=True,
                  norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, init_scale=1.0)
    cases = [
        ("MVD-B (3-D sin-cos table), fc_norm", mvd.VisionTransformer, dict(), 64, 16, 2, "fc_norm", False),
        ("MVD-B + class token (S = 1569), fc_norm", mvd.VisionTransformer, dict(use_cls_token=True), 64, 16, 2, "fc_norm", True),
        ("UMT-B tubelet 1 x 8 frames (S = 1568), fc_norm", umt.VisionTransformer, dict(), 64, 8, 1, "fc_norm", False),
        ("UMT-B tubelet 1 x 16 frames (S = 3136), fc_norm", umt.VisionTransformer, dict(), 24, 16, 1, "fc_norm", False),
        ("ViT-B final_reduction='none' (per-token logits)", mf.VisionTransformer, dict(), 64, 16, 2, "none", False),
        ("ViT-B final_reduction='cls'", mf.VisionTransformer, dict(), 64, 16, 2, "cls", False),
    ]
    for name, cls, extra, B, frames, tubelet, red, cls_tok in cases:
        model = cls(all_frames=frames, tubelet_size=tubelet, final_reduction=red, **common, **extra)
        sd = synth.make_variant_state_dict("vit_base_patch16_224", seed=0, tubelet=tubelet, final_reduction=red,
                                           cls_token=cls_tok)
        if isinstance(model.pos_embed, torch.nn.Parameter):
            sd["pos_embed"] = model.pos_embed.detach().clone()
        model.load_state_dict(sd)
        model = model.to(dev).eval()
        clips = [synth.make_clips(B, seed=90 + i, frames=frames).to(dev).to(torch.bfloat16) for i in range(2)]
        ms = timed(lambda i: model(clips[i % 2]), iters)
        n_patch = (frames // tubelet) * 196
        S = n_patch + (1 if cls_tok else 0)
        gf = (2 * n_patch * 3 * tubelet * 256 * D + L * (24 * S * D * D + 4 * S * S * D)) / 1e9
        cps = B / (ms * 1e-3)
        print(json.dumps({"config": f"{name}, {B} clips/step", "ms_per_step": ms, "clips_per_s": cps, "gflop_per_clip": gf,
                          "model_tflops": cps * gf / 1e3, "frac_of_sustained_peak": cps * gf / 1e3 / peak,
                          "launches": model.prepare().last_launches}), flush=True)
        del model, clips
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--siblings", action="store_true", help="only the MVD / UMT / reduction variants")
    ap.add_argument("--dapt-only", action="store_true", help="only the masked encoder / MAE forward lines")
    ap.add_argument("--small-only", action="store_true", help="of the classifiers, only ViT-S")
    a = ap.parse_args()
    dev = torch.device("cuda")
    peak = 1374.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    if a.siblings:
        return siblings(a.iters, dev, peak)
    classifiers = (("vit_small_patch16_224", 128), ("vit_base_patch16_224", 64), ("vit_large_patch16_224", 32))
    for arch, B in (() if a.dapt_only else classifiers[:1] if a.small_only else classifiers):
        model = mf.__dict__[arch](num_classes=2, all_frames=16, tubelet_size=2, init_scale=1.0, final_reduction="fc_norm")
        model.load_state_dict(synth.make_state_dict(arch, seed=0))
        model = model.to(dev).eval()
        vids = [synth.make_video(B + 15, seed=i).to(dev) for i in range(3)]
        ms = timed(lambda i: model.forward_windows(vids[i % 3], start=0, count=B), a.iters)
        cps = B / (ms * 1e-3)
        print(json.dumps({"config": f"{arch} sliding-window inference, {B} windows/step", "ms_per_step": ms, "clips_per_s": cps,
                          "model_tflops": cps * GF[arch] / 1e3, "frac_of_sustained_peak": cps * GF[arch] / 1e3 / peak}), flush=True)
        del model, vids
        torch.cuda.empty_cache()
    # config 4: DAPT (ViT-B, 90 % tube masking, B = 100 clips)
    arch, B = "vit_base_patch16_224", 100
    full = mp.pretrain_videomae_base_patch16_224(decoder_depth=4)
    full.load_state_dict(synth.make_pretrain_state_dict(arch, seed=6, decoder_depth=4))
    full = full.to(dev).eval()
    clips = [synth.make_clips(B, seed=70 + i).to(dev).to(torch.bfloat16) for i in range(2)]
    gen = TubeMaskingGenerator((8, 14, 14), 0.9)
    masks = [batch_masks(gen, B, dev) for _ in range(2)]
    ms_enc = timed(lambda i: full.encoder(clips[i % 2], masks[i % 2], n_visible=160), a.iters)
    gf_enc = 28.50
    print(json.dumps({"config": f"DAPT masked encoder ViT-B, mask 0.9, {B} clips/step (160 visible tokens)", "ms_per_step": ms_enc,
                      "clips_per_s": B / (ms_enc * 1e-3), "model_tflops": B / (ms_enc * 1e-3) * gf_enc / 1e3}), flush=True)
    ms_full = timed(lambda i: full(clips[i % 2], masks[i % 2], n_visible=160), a.iters)
    N, Dd = 1568, 384
    gf_dec = (4 * (24 * N * Dd * Dd + 4 * N * N * Dd) + 2 * 160 * 768 * Dd + 2 * N * Dd * 1536) / 1e9
    print(json.dumps({"config": f"full MAE pre-training forward ViT-B + 4 decoder blocks, mask 0.9, {B} clips/step",
                      "ms_per_step": ms_full, "clips_per_s": B / (ms_full * 1e-3),
                      "model_tflops": B / (ms_full * 1e-3) * (gf_enc + gf_dec) / 1e3, "gflop_per_clip": gf_enc + gf_dec,
                      "launches": full.prepare().last_launches}), flush=True)


if __name__ == "__main__":
    main()
