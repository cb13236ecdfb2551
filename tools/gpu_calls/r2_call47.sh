#!/bin/bash
O=gpurun_out/r2ah; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attention_kernel' -s 3 -c 1 -f -o $O/attention \
  python tools/profile_step.py --depth 2 --batch 64 --iters 3 > $O/ncu_attention.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
ls -la $O
