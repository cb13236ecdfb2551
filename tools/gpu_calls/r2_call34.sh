#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
for tool in synccheck; do
  echo "=== $tool attention"; timeout 600 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize_small.py attention 2>&1 | grep -v "Host Frame\|=========         in\|=========     Saved" | head -24
done | tee $O/sanitizer_attention_sync.txt
