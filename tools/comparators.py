"""Same-box library comparators (SURVEY §8d "the real bar"): each kernel of libstad.so next to the library kernel the
reference reaches for the same op on this GPU — cuBLASLt (F.linear bf16), flash-attn 2 and cuDNN/SDPA fused attention,
ATen layer_norm — each timed alone with CUDA events, L2 flushed between launches.  Measurement tooling only.
    python tools/comparators.py [--batch 64] [--arch vit_base_patch16_224]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from simple_tad_b200 import _lib as L  # noqa: E402
from tools.bench_kernel import timeit  # noqa: E402

ARCH = {"vit_small_patch16_224": (384, 6), "vit_base_patch16_224": (768, 12), "vit_large_patch16_224": (1024, 16)}


def row(name, ours_ms, lib_ms, flops, lib_name):
    print(f"{name:34s} ours {ours_ms * 1e3:8.1f} us {flops / ours_ms / 1e9:8.1f} TF/s | {lib_name:22s} "
          f"{lib_ms * 1e3:8.1f} us {flops / lib_ms / 1e9:8.1f} TF/s | ours/lib time {ours_ms / lib_ms:5.2f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--arch", default="vit_base_patch16_224")
    ap.add_argument("--tokens", type=int, default=1568)
    a = ap.parse_args()
    D, H = ARCH[a.arch]
    M = a.batch * a.tokens
    dev = "cuda"
    print(f"# {a.arch} batch {a.batch} x {a.tokens} tokens (M = {M}); median of 20, L2 flushed; {torch.cuda.get_device_name()}")
    for name, N, K, mode in (("qkv (LN-fold + bias)", 3 * D, D, "ln"), ("proj (+bias +residual +stats)", D, D, "resid_stats"),
                             ("fc1 (LN-fold + bias + GELU)", 4 * D, D, "ln_gelu"), ("fc2 (+bias +residual +stats)", D, 4 * D, "resid_stats")):
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        bias16 = bias.to(torch.bfloat16)
        if mode == "resid_stats":
            res = torch.randn(M, N, device=dev).to(torch.bfloat16)
            ours = lambda: L.gemm_bias_residual_stats(x, w, bias, res, 1e-6)  # noqa: E731
        else:
            stats = L.row_stats(x, 1e-6)
            cs = w.float().sum(1).contiguous()
            ours = lambda: L.ln_gemm(x, stats, w, bias, cs, gelu=(mode == "ln_gelu"))  # noqa: E731
        o_ms, _ = timeit(ours)
        l_ms, _ = timeit(lambda: F.linear(x, w, bias16))      # the bare cuBLASLt GEMM + bias, none of the fused work
        row(f"{name} {N}x{K}", o_ms, l_ms, 2.0 * M * N * K, "cuBLASLt F.linear+bias")
        del x, w
    # attention
    qkv = torch.randn(a.batch, a.tokens, 3, H, 64, device=dev).to(torch.bfloat16)
    fl = 4.0 * a.batch * H * a.tokens * a.tokens * 64
    o_ms, _ = timeit(lambda: L.attention(qkv))
    try:
        from flash_attn import flash_attn_qkvpacked_func
        fa_ms, _ = timeit(lambda: flash_attn_qkvpacked_func(qkv))
        row(f"attention H={H} S={a.tokens}", o_ms, fa_ms, fl, "flash-attn 2 (qkvpacked)")
    except Exception as e:  # noqa: BLE001
        print("flash_attn unavailable:", repr(e)[:120])
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
    for backend in ("CUDNN_ATTENTION", "FLASH_ATTENTION"):
        try:
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel(getattr(SDPBackend, backend)):
                s_ms, _ = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
            row(f"attention H={H} S={a.tokens}", o_ms, s_ms, fl, f"SDPA {backend.lower()}")
        except Exception as e:  # noqa: BLE001
            print(f"SDPA {backend} unavailable:", repr(e)[:120])
    # LayerNorm statistics: ours come out of the GEMM epilogues (finalize kernel); ATen does a full pass
    x = torch.randn(M, D, device=dev).to(torch.bfloat16)
    g = torch.ones(D, device=dev, dtype=torch.bfloat16)
    ln_ms, _ = timeit(lambda: F.layer_norm(x, (D,), g, g, 1e-6))
    st_ms, _ = timeit(lambda: L.row_stats(x, 1e-6))
    print(f"{'LayerNorm pass over [M, D]':34s} ours (stand-alone stats kernel) {st_ms * 1e3:8.1f} us | ATen layer_norm {ln_ms * 1e3:8.1f} us "
          f"(on the model path neither runs: statistics come from the GEMM epilogues, ~9 us finalize)")


if __name__ == "__main__":
    main()
