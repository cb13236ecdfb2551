#!/bin/bash
# live comparison with the unmodified reference executed on the GPU (fresh seeds, bench batch)
O=gpurun_out/r2au; mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "live_unmodified" -rs 2>&1 | tail -8 | tee $O/live_ref.txt
