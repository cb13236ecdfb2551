#!/bin/bash
# the reference's own GPU stack (autocast fp16 + flash-attn 2 / eager) on this B200 next to the kernels, 64 ViT-B clips per step
O=gpurun_out/r2av; mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -s -k "reference_gpu_stack" -rs 2>&1 | grep -v "^$" | tail -12 | tee $O/ref_gpu_stack.txt
