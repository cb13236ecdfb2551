"""Short workload for ncu: a few forwards of a depth-limited ViT (same widths/kernels as the bench) and a per-shape
table of the library's own per-launch event timings.
    python tools/profile_step.py [--model vit_base_patch16_224] [--depth 2] [--batch 64] [--iters 3] [--table]"""
import argparse
import collections
import os
import sys
from functools import partial

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import synth_data as synth  # noqa: E402  (synthetic weights / inputs)
from simple_tad_b200 import _lib, modeling_finetune as mf  # noqa: E402

EPI = {0: "bias", 1: "ln", 3: "ln+gelu", 4: "resid", 20: "resid+stats", 8: "pos", 40: "patch-embed", 56: "patch-embed+stats", 24: "pos+stats"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vit_base_patch16_224")
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--table", action="store_true")
    a = ap.parse_args()
    D, depth, heads = synth.ARCHS[a.model]
    depth = a.depth or depth
    sd = {k: v for k, v in synth.make_state_dict(a.model, seed=0).items()
          if not k.startswith("blocks.") or int(k.split(".")[1]) < depth}
    model = mf.VisionTransformer(patch_size=16, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16,
                                 tubelet_size=2, init_scale=1.0)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    frames = synth.make_video(a.batch + 15, seed=1).cuda()
    for _ in range(2):
        model.forward_windows(frames, count=a.batch)
    torch.cuda.synchronize()
    if a.table:
        _lib.profile_enable(4096)
    for _ in range(a.iters):
        model.forward_windows(frames, count=a.batch)
    torch.cuda.synchronize()
    if a.table:
        recs = _lib.profile_read(4096)
        _lib.profile_enable(0)
        agg = collections.OrderedDict()
        for kind, epi, m, n, k, ms in recs:
            key = (kind, EPI.get(epi, str(epi)) if kind == "gemm" else "", m, n, k)
            d = agg.setdefault(key, [0, 0.0])
            d[0] += 1
            d[1] += ms
        total = sum(v[1] for v in agg.values())
        print(f"{'kernel':<12}{'epilogue':<12}{'M':>8}{'N':>6}{'K':>6}{'n':>4}{'avg us':>10}{'TFLOP/s':>9}{'GB/s':>8}{'share':>7}")
        for (kind, epi, m, n, k), (cnt, ms) in agg.items():
            avg = ms / cnt
            tf = gb = 0.0
            if kind == "gemm":
                tf = 2.0 * m * n * k / (avg * 1e-3) / 1e12
                gb = 2.0 * (m * k + n * k + m * n) / (avg * 1e-3) / 1e9
            elif kind == "attention":
                tf = 4.0 * m * n * float(k) * k * 64 / (avg * 1e-3) / 1e12
            elif kind in ("row_stats", "pool", "layernorm"):
                gb = 2.0 * m * n / (avg * 1e-3) / 1e9
            print(f"{kind:<12}{epi:<12}{m:>8}{n:>6}{k:>6}{cnt:>4}{avg * 1e3:>10.1f}{tf:>9.1f}{gb:>8.0f}{100 * ms / total:>6.1f}%")


if __name__ == "__main__":
    main()
