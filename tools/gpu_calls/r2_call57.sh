#!/bin/bash
# edge-shape checks (attention sequence lengths at the seams of the tile geometry, GEMM shapes of one row / one k-block / N = 320)
O=gpurun_out/r2aq; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "edges" 2>&1 | tail -25 | tee $O/edges.txt
