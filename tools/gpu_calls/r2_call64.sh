#!/bin/bash
# full GPU suite with the live-reference tests, as the driver runs it
O=gpurun_out/r2ax; mkdir -p $O
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee $O/gpu_tests.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a $O/gpu_tests.txt
