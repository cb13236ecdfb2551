#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 20 -c 5 -f -o $O/vits \
  python tools/profile_step.py --model vit_small_patch16_224 --depth 2 --batch 128 --iters 3 > $O/ncu_vits.log 2>&1
ls -la $O
