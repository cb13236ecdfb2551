#!/bin/bash
O=gpurun_out/r2ak; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -15 | tee $O/att_tests.txt
