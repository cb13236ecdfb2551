#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
ATT_SHAPE=64,12,1568 STAD_LIB=build_variants/libstad_trace.so timeout 120 python tools/att_trace.py 0 200000 > $O/trace_s1568.txt 2>&1
wc -l $O/trace_s1568.txt
