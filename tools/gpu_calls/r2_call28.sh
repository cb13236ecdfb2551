#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
for rep in 1 2; do
for lib in simple-tad_b200/libstad.so build_variants/libstad_testfirst.so; do
  STAD_LIB=$lib timeout 120 python tools/bench_kernel.py attention 64 12 1568 2>&1 | tail -1
  STAD_LIB=$lib timeout 120 python tools/bench_kernel.py attention 100 12 160 2>&1 | tail -1
done; done | tee $O/ab_testfirst.txt
ATT_SHAPE=64,12,1568 STAD_LIB=build_variants/libstad_testfirst_trace.so timeout 120 python tools/att_trace.py 0 200000 > $O/trace_s1568_testfirst.txt 2>&1
