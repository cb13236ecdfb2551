#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -2 | tee $O/kernel_tests.txt
grep -q failed $O/kernel_tests.txt && exit 1
for tool in racecheck synccheck memcheck; do
  echo "=== $tool attention"; timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py attention 2>&1 | grep -v "Host Frame\|=========         in" | tail -8
done | tee $O/sanitizer_attention.txt
for s in "64 12 1568" "100 12 160"; do timeout 120 python tools/bench_kernel.py attention $s 2>&1 | tail -1; done | tee $O/att_times.txt
