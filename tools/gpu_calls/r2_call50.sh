#!/bin/bash
O=gpurun_out/r2aj; mkdir -p $O
timeout 120 python tools/bench_kernel.py attention 64 12 1568 2>&1 | tail -1
timeout 300 python tools/comparators.py 2>&1 | grep attention
timeout 120 python tools/bench_kernel.py attention 64 12 1568 2>&1 | tail -1
timeout 120 python - <<'P'
import sys; sys.path.insert(0,'.')
import torch
from simple_tad_b200 import _lib as L
from tools.bench_kernel import timeit
for trial in range(3):
    qkv = torch.randn(64, 1568, 3, 12, 64, device="cuda").to(torch.bfloat16)
    print("fresh tensor", trial, timeit(lambda: L.attention(qkv)))
x = torch.randn(100352, 768, device="cuda").to(torch.bfloat16)
w = torch.randn(2304, 768, device="cuda").to(torch.bfloat16)
for _ in range(200): torch.nn.functional.linear(x, w)
torch.cuda.synchronize()
print("after 200 GEMMs", timeit(lambda: L.attention(qkv)))
P
