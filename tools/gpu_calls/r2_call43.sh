#!/bin/bash
O=gpurun_out/r2ae; mkdir -p $O
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "mae_vitb_full" 2>&1 | tail -30 | tee $O/mae_fail.txt
