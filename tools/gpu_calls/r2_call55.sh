#!/bin/bash
# compute-sanitizer over the final library: whole small forward (folded statistics with prefetched partials, patch embed,
# pool / head, masked encoder), single-CTA and CTA-pair GEMM tiles, row kernels
O=gpurun_out/r2ao; mkdir -p $O
for tool in racecheck synccheck memcheck; do
  echo "=== $tool model gemm pair rows"; timeout 900 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize_small.py model gemm pair rows 2>&1 | grep -v "Host Frame\|=========         in\|=========     Saved" | head -30
done | tee $O/sanitizer_final.txt
