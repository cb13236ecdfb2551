#!/bin/bash
O=gpurun_out/r2ag; mkdir -p $O
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "two_ranks_nccl" 2>&1 | tail -60 > $O/nccl_test_full.txt
grep -n "Error\|error\|assert\|Exception" $O/nccl_test_full.txt | head -20
