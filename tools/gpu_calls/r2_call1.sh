#!/bin/bash
# Round-2 GPU call 1: attention hand-over experiments (STAD_ATT_NO_MMA etc.), HBM-bound kernel evidence (event timings
# + ncu DRAM traffic), compute-sanitizer logs, library comparators.  Everything lands in gpurun_out/r2a/.
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.txt 2>&1
for v in "" nomma poly0 poly2; do
  if [ -z "$v" ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  for i in 1 2; do STAD_LIB=$lib timeout 300 python tools/bench_kernel.py attention 64 12 1568; done
done > $O/att_variants.txt 2>&1
STAD_LIB=build_variants/libstad_trace.so timeout 300 python tools/att_trace.py 200000 215000 > $O/trace_base.txt 2>&1
STAD_LIB=build_variants/libstad_nomma_trace.so timeout 300 python tools/att_trace.py 200000 215000 > $O/trace_nomma.txt 2>&1
timeout 600 python tools/hbm_kernels.py --json $O/hbm_kernels.json > $O/hbm_kernels.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'pool_partial|pool_head|cast_kernel|normalize_u8|row_norm|gather_patches|decoder_assemble|tail_rows' \
  -o $O/hbm python tools/hbm_kernels.py --once > $O/ncu_hbm.log 2>&1
for tool in racecheck synccheck memcheck; do
  for grp in gemm pair attention rows; do
    echo "=== $tool $grp"; timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py $grp 2>&1 | tail -25
  done
done > $O/sanitizer.txt 2>&1
timeout 600 python tools/comparators.py > $O/comparators.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.txt 2>&1
tail -3 $O/tests.txt; cat $O/att_variants.txt; tail -12 $O/hbm_kernels.txt
