#!/bin/bash
O=gpurun_out/r2af; mkdir -p $O
for rep in 1 2; do
for v in ship foldsmall; do
  if [ $v = ship ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  STAD_LIB=$lib timeout 600 python bench.py --no-extras --no-cpu-baseline > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err
  python -c "
import json
d=json.loads(open('$O/bench_${v}_$rep.json').read())
print('$v', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['attention']['avg_launch_ms'], d['clocks']['sm_mhz'])"
done; done | tee $O/ab_fold.txt
