#!/bin/bash
O=gpurun_out/r2aa; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -3 | tee $O/kernel_tests.txt
grep -q failed $O/kernel_tests.txt && exit 1
for s in "64 12 1568" "100 12 160" "1 12 1568"; do timeout 120 python tools/bench_kernel.py attention $s 2>&1 | tail -1; done | tee $O/att_times.txt
timeout 300 python tools/power_probe.py 2.5 2>&1 | tee $O/power_probe.txt
timeout 900 python bench.py --no-extras --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['attention'], d.get('clocks'))"
