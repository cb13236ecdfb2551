#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
ATT_SHAPE=100,12,160 STAD_LIB=build_variants/libstad_trace.so timeout 120 python tools/att_trace.py 0 60000 > $O/trace_s160.txt 2>&1
wc -l $O/trace_s160.txt; tail -3 $O/trace_s160.txt
