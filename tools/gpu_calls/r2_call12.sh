#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.txt 2>&1; tail -25 $O/tests.txt
timeout 600 python tools/hbm_kernels.py --json $O/hbm_kernels.json 2>&1 | tee $O/hbm_kernels.txt
