#!/bin/bash
O=gpurun_out/r2ac; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -2 | tee $O/kernel_tests.txt
grep -q failed $O/kernel_tests.txt && exit 1
for rep in 1 2; do
for lib in simple-tad_b200/libstad.so build_variants/libstad_noobs.so build_variants/libstad_prewait.so; do
  STAD_LIB=$lib timeout 120 python tools/bench_kernel.py attention 64 12 1568 2>&1 | tail -1
done; done | tee $O/ab.txt
for tool in synccheck; do
  echo "=== $tool attention"; timeout 600 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize_small.py attention 2>&1 | grep -v "Host Frame\|=========         in\|=========     Saved" | head -12
done | tee $O/sanitizer_attention.txt
