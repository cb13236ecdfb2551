#!/bin/bash
# Evidence pass (round 2): ncu launch list of the bench command, ncu --set full of one block's GEMMs + attention, and of
# the HBM-bound kernels.  A number printed under ncu is never a bench value.
O=gpurun_out/r2n; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_steps2.csv \
  python bench.py --steps 2 --warmup 3 --preheat-s 0 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 40 -c 12 -f -o $O/block \
  python tools/profile_step.py --depth 2 --batch 64 --iters 3 > $O/ncu_block.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'pool_partial|pool_head|cast_kernel|normalize_u8|row_norm|gather_patches|decoder_assemble|tail_rows|stats_finalize' \
  -f -o $O/hbm python tools/hbm_kernels.py --once > $O/ncu_hbm.log 2>&1
ls -la $O
