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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    peak = 1374.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    for arch, B in (("vit_small_patch16_224", 128), ("vit_base_patch16_224", 64), ("vit_large_patch16_224", 32)):
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
