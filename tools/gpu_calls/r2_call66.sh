#!/bin/bash
# the 2-rank NCCL test of the GPU suite on a 2-GPU box (skipped on one GPU)
O=gpurun_out/r2az; mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "two_ranks_nccl" -rs 2>&1 | tail -4 | tee $O/nccl_test.txt
