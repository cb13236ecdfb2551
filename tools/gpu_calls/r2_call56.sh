#!/bin/bash
# stats_finalize with all partials of a row in flight (ship) against the one-load-per-iteration loop (oldfin): tests, then in-step A/B
O=gpurun_out/r2ap; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3 | tee $O/tests.txt
grep -q failed $O/tests.txt && exit 1
for rep in 1 2 3; do
for v in oldfin ship; do
  if [ $v = ship ]; then lib=simple-tad_b200/libstad.so; else lib=build_variants/libstad_$v.so; fi
  STAD_LIB=$lib timeout 600 python bench.py --no-extras --no-cpu-baseline > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err
  python -c "
import json
d=json.loads(open('$O/bench_${v}_$rep.json').read())
print('$v', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['attention']['avg_launch_ms'], d['clocks']['sm_mhz'], 'row_stats share', d['roofline']['other_share_of_step'].get('row_stats'), 'kernel ms', d['roofline']['kernel_ms_per_step'])"
done; done | tee $O/finalize_ab.txt
