#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3
for s in "64 12 1568" "100 12 160" "64 12 1569" "1 12 1568" "128 6 1568"; do timeout 120 python tools/bench_kernel.py attention $s; done 2>&1 | tee $O/att_times.txt
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "config3 or config5 or config4" 2>&1 | tail -3
timeout 300 python tools/profile_dapt.py 100 > $O/dapt.txt 2>&1; head -12 $O/dapt.txt
timeout 900 python bench.py --no-extras --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['attention'])"
