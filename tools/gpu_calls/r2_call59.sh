#!/bin/bash
# pool_partial with 384-thread blocks (whole row groups, no idle threads) and eight rows in flight (ship) against 256 threads / four rows (oldpool)
O=gpurun_out/r2as; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "pool or rows_norm" 2>&1 | tail -2 | tee $O/tests.txt
for rep in 1 2; do for v in oldpool ship; do
  if [ $v = ship ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  echo "== $v"; STAD_LIB=$lib timeout 300 python tools/hbm_kernels.py --only pool 2>&1 | tail -2
done; done | tee $O/pool_ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pool_partial -c 3 python tools/hbm_kernels.py --only pool --iters 2 2>&1 | grep -E "pool_partial|gpu__time|dram__bytes" | tail -12 | tee $O/pool_ncu.txt
