#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > $O/tests_att.txt 2>&1; tail -3 $O/tests_att.txt
for i in 1 2; do STAD_LIB=simple-tad_b200/libstad.so timeout 300 python tools/bench_kernel.py attention 64 12 1568; done 2>&1 | tee $O/att.txt
STAD_LIB=build_variants/libstad_trace2.so timeout 300 python tools/att_trace.py 200000 212000 > $O/trace_new.txt 2>&1
timeout 600 python tools/comparators.py 2>&1 | grep attention | tee $O/comparators_att.txt
