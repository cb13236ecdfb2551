#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
for lib in simple-tad_b200/libstad.so build_variants/libstad_ofulltry.so; do
  echo "=== synccheck attention $lib"; STAD_LIB=$lib timeout 600 compute-sanitizer --tool synccheck --print-limit 3 python tools/sanitize_small.py attention 2>&1 | grep -v "Host Frame\|=========         in\|=========     Saved" | head -12
  STAD_LIB=$lib timeout 120 python tools/bench_kernel.py attention 64 12 1568 2>&1 | tail -1
  STAD_LIB=$lib timeout 120 python tools/bench_kernel.py attention 100 12 160 2>&1 | tail -1
done | tee $O/sanitizer_attention_sync.txt
