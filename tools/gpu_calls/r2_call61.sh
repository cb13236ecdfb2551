#!/bin/bash
# live comparison with the unmodified reference executed on the GPU (fresh seeds, bench batch; classifier, DAPT encoder, MAE)
O=gpurun_out/r2au; mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -s -k "live_unmodified" -rs 2>&1 | grep -v "^$" | tail -14 | tee $O/live_ref.txt
