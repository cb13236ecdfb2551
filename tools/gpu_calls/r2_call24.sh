#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3 | tee $O/att_tests.txt
grep -q failed $O/att_tests.txt && exit 1
for s in "100 12 160" "64 12 1568" "64 12 1569" "1 12 1568" "8 12 1568" "100 6 1568" "256 12 160" "64 12 392"; do timeout 120 python tools/bench_kernel.py attention $s 2>&1 | tail -1; done | tee $O/att_times.txt
grep -q Error $O/att_times.txt && exit 1
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_small.py attention 2>&1 | tail -5 | tee $O/racecheck_att.txt
timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -3 | tee $O/model_tests.txt
timeout 300 python tools/profile_dapt.py 100 > $O/dapt.txt 2>&1; head -12 $O/dapt.txt
timeout 900 python bench.py --no-extras --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['attention'])"
