#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
for v in pf0 new; do
  if [ $v = new ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  STAD_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -f -o $O/att_$v python tools/bench_kernel.py attention 64 12 1568 > $O/ncu_$v.log 2>&1
done
ls -la $O
