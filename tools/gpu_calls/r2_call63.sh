#!/bin/bash
# the reference's GPU stack on the DAPT encoder (fp16 autocast + flash-attn 2) next to this repo's encoder, 100 clips, mask 0.9
O=gpurun_out/r2aw; mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -s -k "dapt_encoder_faster" -rs 2>&1 | grep -v "^$" | tail -8 | tee $O/ref_gpu_dapt.txt
