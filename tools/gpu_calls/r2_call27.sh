#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3 | tee $O/att_tests.txt
grep -q failed $O/att_tests.txt && exit 1
for s in "64 12 1568" "100 12 160" "64 12 1569" "1 12 1568" "8 12 1568" "256 12 160" "64 12 392"; do timeout 120 python tools/bench_kernel.py attention $s 2>&1 | tail -1; done | tee $O/att_times.txt
grep -q Error $O/att_times.txt && exit 1
timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -3 | tee $O/model_tests.txt
