#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 40 -c 6 -f -o $O/block \
  python tools/profile_step.py --depth 2 --batch 64 --iters 3 > $O/ncu_block.log 2>&1
ls -la $O
