"""Small-shape workload for compute-sanitizer (racecheck / synccheck / memcheck) over the mbarrier-heavy kernels:
single-CTA and CTA-pair GEMM tiles with every epilogue, and the attention kernel with two-slot, one-slot, ragged and
key-split units.  Results are still checked against PyTorch fp32 (tests/kernel_checks.py), so a sanitizer-clean run
is also a correct run.
    compute-sanitizer --tool racecheck python tools/sanitize_small.py [gemm|pair|attention|rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import kernel_checks as kc  # noqa: E402

GROUPS = {
    "gemm": lambda: [kc.check_gemm(300, 192, 128, "bias", seed=1), kc.check_gemm(256, 256, 256, "resid", seed=2),
                     kc.check_gemm(200, 384, 384, "ln", seed=3), kc.check_gemm(257, 512, 128, "ln_gelu", seed=4),
                     kc.check_gemm_stats(384, 384, 256, 128, seed=5)],
    # pair tiles need N % 256 == 0 and >= 74 pair tiles: 9472 x 512 -> 37 x 2 pair tiles; K kept small
    "pair": lambda: [kc.check_gemm(9472, 512, 2048, "bias", seed=6), kc.check_gemm(9600, 768, 768, "resid", seed=7),
                     kc.check_gemm(9472, 1024, 768, "ln_gelu", seed=8), kc.check_gemm_stats(9472, 768, 512, 768, seed=9)],
    "attention": lambda: [kc.check_attention(1, 2, 288, seed=1), kc.check_attention(1, 1, 160, seed=2),
                          kc.check_attention(2, 1, 100, seed=3), kc.check_attention(1, 1, 417, peaky=4.0, seed=4),
                          kc.check_attention(40, 4, 160, seed=5), kc.check_attention_outliers(1, 2, 288, seed=6),
                          kc.check_attention_outliers(20, 4, 392, seed=7)],
    "rows": lambda: [kc.check_pool_head(3, 160, 384), kc.check_layernorm(100, 768), kc.check_row_stats(100, 384),
                     kc.check_normalize_u8(2, 64, 64), kc.check_decoder_assemble(2, 196, 20, 192, seed=50)],
}


def _model_forward():
    """Whole small forward (ViT-S width, 2 blocks, 2 clips = 3136 rows): the path on which the LayerNorm-folded GEMMs finish
    the statistics themselves from the prefetched partial sums (M <= 32768), the patch-embed GEMM with its 5-D tensor map,
    pool + head; plus the masked encoder (visible-token gather).  Two runs must agree bit for bit."""
    from functools import partial
    import synth_data as synth
    from simple_tad_b200 import modeling_finetune as mf, modeling_pretrain as mp
    arch = "vit_small_d2"
    D, depth, heads = synth.ARCHS[arch]
    model = mf.VisionTransformer(patch_size=16, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16,
                                 tubelet_size=2, init_scale=1.0, final_reduction="fc_norm")
    model.load_state_dict(synth.make_state_dict(arch, seed=7))
    model = model.cuda().eval()
    x = synth.make_clips(2, seed=7).cuda()
    a = model(x)
    b = model(x)
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and torch.equal(a, b)
    enc = mp.PretrainVisionTransformerEncoder(embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.)
    enc.load_state_dict(synth.make_state_dict(arch, seed=8, encoder=True))
    enc = enc.cuda().eval()
    y = enc(x, synth.tube_mask(2, 0.9, seed=8).cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    return [a, y]


GROUPS["model"] = _model_forward


def main():
    names = sys.argv[1:] or list(GROUPS)
    for n in names:
        GROUPS[n]()
        torch.cuda.synchronize()
        print(f"sanitize_small: {n} ok", flush=True)


if __name__ == "__main__":
    main()
