#!/bin/bash
# full GPU suite (with the new full-size property checks) + smoke; A/B of the unrolled stats_finalize inside the bench step
O=gpurun_out/r2an; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/gpu_tests.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $O/smoke.txt
for rep in 1 2; do
for v in ship oldfin; do
  if [ $v = ship ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  STAD_LIB=$lib timeout 600 python bench.py --no-extras --no-cpu-baseline > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err
  python -c "
import json
d=json.loads(open('$O/bench_${v}_$rep.json').read())
print('$v', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['attention']['avg_launch_ms'], d['clocks']['sm_mhz'], 'row_stats share', d['roofline']['other_share_of_step'].get('row_stats'), 'kernel ms', d['roofline']['kernel_ms_per_step'])"
done; done | tee $O/finalize_ab.txt
