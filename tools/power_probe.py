"""Sustained clock / power of each hot kernel on its own: every kernel is launched back to back for ~2.5 s while NVML is
sampled every 20 ms; the last second gives the settled SM clock, board power and the kernel's time at that clock.
Tells which kernels are limited by the 1 kW power cap inside a step and which by their own chain.
    python tools/power_probe.py [seconds]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pynvml  # noqa: E402
import torch  # noqa: E402

from simple_tad_b200 import _lib as L  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self, handle):
        super().__init__(daemon=True)
        self.h, self.samples, self.stop_flag = handle, [], False

    def run(self):
        while not self.stop_flag:
            self.samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM),
                                 pynvml.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            time.sleep(0.02)


def probe(name, fn, flops, handle, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = Sampler(handle)
    s.start()
    t0 = time.time()
    n_tail, e0, e1 = 0, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    started = False
    while True:
        el = time.time() - t0
        if el > seconds:
            break
        if not started and el > seconds - 1.0:
            e0.record()
            started = True
        for _ in range(20):
            fn()
        if started:
            n_tail += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop_flag = True
    s.join()
    tail = [x for x in s.samples if x[0] - t0 > seconds - 1.0]
    clk = sorted(x[1] for x in tail)[len(tail) // 2]
    pw = sorted(x[2] for x in tail)[len(tail) // 2]
    us = e0.elapsed_time(e1) * 1e3 / max(n_tail, 1)
    print(f"{name:34s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  SM {clk:5d} MHz  {pw:6.1f} W  "
          f"({us * clk / 1e3:8.1f} kclk)", flush=True)
    time.sleep(1.0)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 2.5
    quick = len(sys.argv) > 2 and sys.argv[2] in ("quick", "attention")  # no cuBLAS comparators
    only_attention = len(sys.argv) > 2 and sys.argv[2] == "attention"
    pynvml.nvmlInit()
    handle = pynvml.nvmlDeviceGetHandleByIndex(0)
    B, S, D, H = 64, 1568, 768, 12
    M = B * S
    qkv = torch.randn(B, S, 3, H, 64, device="cuda").to(torch.bfloat16)
    probe("attention 64x12x1568", lambda: L.attention(qkv), 4.0 * B * H * S * S * 64, handle, seconds)
    if only_attention:
        return
    a = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    ah = torch.randn(M, 4 * D, device="cuda").to(torch.bfloat16)
    res = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    stats = L.row_stats(a, 1e-6)
    for name, N, K, mode in (("qkv  ln", 3 * D, D, "ln"), ("fc1  ln+gelu", 4 * D, D, "ln_gelu"),
                             ("proj resid+stats", D, D, "rs"), ("fc2  resid+stats", D, 4 * D, "rs")):
        w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        x = ah if K == 4 * D else a
        if mode == "rs":
            fn = lambda x=x, w=w, bias=bias: L.gemm_bias_residual_stats(x, w, bias, res, 1e-6)  # noqa: E731
        else:
            colsum = w.float().sum(1).contiguous()
            fn = lambda x=x, w=w, bias=bias, colsum=colsum, mode=mode: L.ln_gemm(  # noqa: E731
                x, stats, w, bias, colsum, gelu=(mode == "ln_gelu"))
        probe(f"gemm {name} [{M}x{N}x{K}]", fn, 2.0 * M * N * K, handle, seconds)
        if quick:
            continue
        wt = w.t().contiguous()
        probe(f"  torch.matmul (cuBLAS) same shape", lambda x=x, wt=wt: torch.matmul(x, wt), 2.0 * M * N * K, handle, seconds)


if __name__ == "__main__":
    main()
