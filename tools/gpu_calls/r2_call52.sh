#!/bin/bash
# in-step A/B of the share of polynomial exp2 in the attention softmax (energy-bound step: does all-MUFU cost less energy?)
O=gpurun_out/r2al; mkdir -p $O
for v in ship poly0 poly6 poly8; do
  if [ $v = ship ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  echo "=== $v"
  STAD_LIB=$lib timeout 120 python tools/bench_kernel.py attention 64 12 1568 2>&1 | tail -1
  STAD_LIB=$lib timeout 300 python tools/power_probe.py 2.5 attention 2>&1 | tail -1
done | tee $O/poly_ab.txt
for rep in 1 2; do
for v in ship poly0 poly6 poly8; do
  if [ $v = ship ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  STAD_LIB=$lib timeout 600 python bench.py --no-extras --no-cpu-baseline > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err
  python -c "
import json
d=json.loads(open('$O/bench_${v}_$rep.json').read())
print('$v', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['attention']['avg_launch_ms'], d['clocks']['sm_mhz'])"
done; done | tee $O/poly_bench_ab.txt
