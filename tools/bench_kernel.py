"""Time one kernel of libstad.so in isolation (CUDA events, L2 flushed between launches).
    python tools/bench_kernel.py attention [B H S]
    python tools/bench_kernel.py gemm M N K mode      (mode: plain | bias | resid | ln | ln_gelu)
STAD_LIB=<path> selects a development build of the library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from simple_tad_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=20, warmup=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def main():
    kind = sys.argv[1]
    if kind == "attention":
        B, H, S = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (64, 12, 1568)
        qkv = (torch.randn(B, S, 3, H, 64, device="cuda") * 1.0).to(torch.bfloat16)
        med, best = timeit(lambda: L.attention(qkv))
        fl = 4.0 * B * H * S * S * 64
        print(f"attention B={B} H={H} S={S}: median {med * 1e3:.1f} us ({fl / med / 1e9:.1f} TFLOP/s), "
              f"best {best * 1e3:.1f} us ({fl / best / 1e9:.1f} TFLOP/s)  lib={L.LIB_PATH}")
    elif kind == "gemm":
        M, N, K = (int(x) for x in sys.argv[2:5])
        mode = sys.argv[5] if len(sys.argv) > 5 else "plain"
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        if mode in ("plain", "bias", "resid"):
            res = torch.randn(M, N, device="cuda").to(torch.bfloat16) if mode == "resid" else None
            fn = lambda: L.gemm_bias_residual(a, w, None if mode == "plain" else bias, res)  # noqa: E731
        elif mode == "resid_stats":
            res = torch.randn(M, N, device="cuda").to(torch.bfloat16)
            fn = lambda: L.gemm_bias_residual_stats(a, w, bias, res, 1e-6)  # noqa: E731
        else:
            stats = L.row_stats(a, 1e-6)
            colsum = w.float().sum(1).contiguous()
            fn = lambda: L.ln_gemm(a, stats, w, bias, colsum, gelu=(mode == "ln_gelu"))  # noqa: E731
        med, best = timeit(fn)
        fl = 2.0 * M * N * K
        print(f"gemm {mode} M={M} N={N} K={K}: median {med * 1e3:.1f} us ({fl / med / 1e9:.1f} TFLOP/s), "
              f"best {best * 1e3:.1f} us ({fl / best / 1e9:.1f} TFLOP/s)")


if __name__ == "__main__":
    main()
