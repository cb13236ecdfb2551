#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
for rep in 1 2 3; do
for lib in simple-tad_b200/libstad.so build_variants/libstad_noobs.so; do
  STAD_LIB=$lib timeout 120 python tools/bench_kernel.py attention 64 12 1568 2>&1 | tail -1
done; done | tee $O/ab_ofull_observe.txt
