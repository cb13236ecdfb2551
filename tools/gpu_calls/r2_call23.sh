#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3 | tee $O/att_tests.txt
grep -q failed $O/att_tests.txt && { timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "
import sys; sys.path.insert(0,'.')
from tests import kernel_checks as kc
import torch
kc.check_attention(5, 12, 1568, seed=3); torch.cuda.synchronize(); print('a ok', flush=True)
kc.check_attention(40, 12, 160, seed=4); torch.cuda.synchronize(); print('b ok', flush=True)
kc.check_attention(16, 6, 392, peaky=5.0, seed=5); torch.cuda.synchronize(); print('c ok', flush=True)
" 2>&1 | grep -v "Host Frame\|=========         in" | head -50; exit 1; }
for s in "64 12 1568" "100 12 160" "64 12 1569" "1 12 1568" "128 6 1568"; do timeout 120 python tools/bench_kernel.py attention $s; done 2>&1 | tee $O/att_times.txt
