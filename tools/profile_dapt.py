"""Per-kernel device times (stad_profile_*: CUDA events around every launch) of one DAPT masked-encoder forward and one
full MAE pre-training forward (ViT-B, mask 0.9, B clips).   python tools/profile_dapt.py [B]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import synth_data as synth  # noqa: E402
from simple_tad_b200 import _lib as L, modeling_pretrain as mp  # noqa: E402
from simple_tad_b200.masking_generator import TubeMaskingGenerator, batch_masks  # noqa: E402


def report(title, recs, skip):
    agg = collections.OrderedDict()
    for kind, epi, m, n, k, ms in recs[skip:]:
        key = (kind, epi, m, n, k)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"# {title}: {total:.3f} ms in {sum(v[0] for v in agg.values())} launches")
    for (kind, epi, m, n, k), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{kind:10s} epi={epi:3d} [{m} x {n} x {k}]  x{cnt:3d}  {ms * 1e3 / cnt:8.1f} us each  {ms:7.3f} ms  {100 * ms / total:5.1f} %")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    dev = torch.device("cuda")
    arch = "vit_base_patch16_224"
    full = mp.pretrain_videomae_base_patch16_224(decoder_depth=4)
    full.load_state_dict(synth.make_pretrain_state_dict(arch, seed=6, decoder_depth=4))
    full = full.to(dev).eval()
    clips = synth.make_clips(B, seed=70).to(dev).to(torch.bfloat16)
    mask = batch_masks(TubeMaskingGenerator((8, 14, 14), 0.9), B, dev)
    for name, fn in (("DAPT masked encoder", lambda: full.encoder(clips, mask, n_visible=160)),
                     ("full MAE forward", lambda: full(clips, mask, n_visible=160))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        L.profile_enable(4096)
        fn()
        torch.cuda.synchronize()
        recs = L.profile_read()
        L.profile_enable(0)
        report(f"{name}, ViT-B, mask 0.9, B={B}", recs, 0)


if __name__ == "__main__":
    main()
