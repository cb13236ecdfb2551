#!/bin/bash
# tools/ab_attention.sh <out file> <variant names...>: isolated attention timing of build_variants/libstad_<v>.so ("ship" = the
# shipped library), two interleaved rounds.
out=$1; shift
for i in 1 2; do
  for v in "$@"; do
    if [ "$v" = ship ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
    STAD_LIB=$lib timeout 300 python tools/bench_kernel.py attention 64 12 1568
  done
done 2>&1 | tee $out
