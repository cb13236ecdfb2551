"""HBM-bound kernels of libstad.so, each alone at the shapes of the bench step (ViT-B, 64 clips) / the DAPT step:
CUDA-event time (L2 flushed between launches) -> achieved GB/s of the ALGORITHMIC bytes against the measured copy
bandwidth (MEASURED_PEAKS.json).  Run it under `ncu --set full -k regex:<kernel>` for the DRAM traffic of one launch.
    python tools/hbm_kernels.py [--only pool,cast,...] [--iters 20] [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from simple_tad_b200 import _lib as L  # noqa: E402
from tools.bench_kernel import timeit  # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


def cases(B=64, N=1568, D=768):
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)
    M = B * N
    out = {}

    x = torch.randn(B, N, D, generator=g).to(torch.bfloat16).to(dev)
    gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    wh, bh = torch.randn(2, D, device=dev) * 0.05, torch.zeros(2, device=dev)
    out["pool"] = (lambda: L.pool_norm_head(x, gam, bet, wh, bh, 1e-6, want_probs=True), 2.0 * M * D,
                   "pool_partial + pool_head: mean(1) -> fc_norm -> head (mf:325-334); 2 B N D read")
    frames = torch.randn(B + 15, 3, 224, 224, generator=g).to(dev)
    out["cast"] = (lambda: L.cast_f32_bf16(frames), 6.0 * frames.numel(), "cast_kernel fp32 -> bf16 (eff:428); 4n + 2n")
    u8 = torch.randint(0, 256, (B + 15, 224, 224, 3), dtype=torch.uint8, generator=g).to(dev)
    out["normalize"] = (lambda: L.normalize_frames_u8(u8, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), bgr=True),
                        9.0 * (B + 15) * 224 * 224, "normalize_u8_kernel (ri:15-34); 3 F H W read + 6 F H W written")
    x2 = x.view(M, D)
    out["row_stats"] = (lambda: L.row_stats(x2, 1e-6), 2.0 * M * D + 8.0 * M, "row_norm_kernel<false>; 2 M D + 8 M")
    # encoder norm of the DAPT batch (100 clips x 160 visible tokens) and of a bench-sized stream
    xe = torch.randn(100 * 160, D, generator=g).to(torch.bfloat16).to(dev)
    out["layernorm_dapt"] = (lambda: L.layernorm(xe, gam, bet, 1e-6), 6.0 * xe.numel(),
                             "row_norm_kernel<true> on [16000, 768] (mp:107); 2 M D read + 4 M D written")
    out["layernorm"] = (lambda: L.layernorm(x2, gam, bet, 1e-6), 6.0 * M * D,
                        "row_norm_kernel<true> on [100352, 768]; 2 M D read + 4 M D written")
    # visible-token gather of the DAPT batch: 100 clips, 160 of 1568 tokens, K = 1536
    clips = torch.randn(100, 3, 16, 224, 224, generator=g).to(torch.bfloat16).to(dev)
    idx = torch.stack([torch.randperm(1568, generator=g)[:160].sort().values for _ in range(100)]).to(torch.int32).to(dev)
    wp = (torch.randn(D, 1536, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    pos = torch.zeros(1568, D, device=dev)
    dims = L.make_dims(dim=D)
    out["gather_embed"] = (lambda: L.patch_embed(clips, wp, pos, dims, 100, 160, tok_idx=idx), None,
                           "gather_patches_kernel (mp:98) followed by the embed GEMM (time is both launches)")
    # decoder assemble: 100 clips, Dd = 384
    vis = torch.randn(100, 160, 384, generator=g).to(torch.bfloat16).to(dev)
    posd = torch.randn(1568, 384, device=dev)
    mt = torch.randn(384, device=dev)
    midx = torch.stack([torch.randperm(1568, generator=g)[:1408].sort().values for _ in range(100)]).to(torch.int32).to(dev)
    out["assemble"] = (lambda: L.decoder_assemble(vis, posd, mt, midx, 1568, 1e-6),
                       2.0 * vis.numel() + 2.0 * 100 * 1568 * 384 + 8.0 * 100 * 1568,
                       "decoder_assemble_kernel (mp:283-288); 2 B n_vis Dd read + 2 B N Dd + 8 B N written")
    xt = torch.randn(100, 1568, 1536, generator=g).to(torch.bfloat16).to(dev)
    out["tail_rows"] = (lambda: L.tail_rows_f32(xt, 1408), 6.0 * 100 * 1408 * 1536,
                        "tail_rows_f32_kernel (mp:174); 2 + 4 bytes per kept element")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default="")
    ap.add_argument("--once", action="store_true", help="one untimed launch per kernel (for ncu captures)")
    a = ap.parse_args()
    peak, how = peak_gbs()
    sel = [s for s in a.only.split(",") if s]
    rows = []
    L.profile_enable(1 << 12)
    for name, (fn, nbytes, what) in cases().items():
        if sel and name not in sel:
            continue
        if a.once:
            fn()
            torch.cuda.synchronize()
            continue
        L.profile_read()
        med, best = timeit(fn, iters=a.iters)
        recs = L.profile_read()
        # the library's own per-launch events split multi-kernel entry points (gather + GEMM, pool pair)
        per_kind = {}
        n_calls = a.iters + 3
        for kind, _epi, _m, _n, _k, ms in recs:
            per_kind.setdefault(kind, []).append(ms)
        kinds = {k: sorted(v)[len(v) // 2] * (len(v) // n_calls or 1) for k, v in per_kind.items()}
        row = {"kernel": name, "what": what, "median_us": med * 1e3, "best_us": best * 1e3, "algorithmic_bytes": nbytes,
               "per_kind_median_us": {k: v * 1e3 for k, v in kinds.items()}}
        if nbytes:
            row["achieved_gbs"] = nbytes / med / 1e6
            row["frac_of_%s_peak" % how] = row["achieved_gbs"] / peak
        rows.append(row)
        gbs = f"{row['achieved_gbs']:8.1f} GB/s = {row['achieved_gbs'] / peak:5.2f} of {how} {peak:.0f}" if nbytes else " " * 20
        print(f"{name:16s} median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {gbs}   {kinds}", flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
