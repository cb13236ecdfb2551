#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" 2>&1 | tail -3
timeout 420 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_small.py pair 2>&1 | tail -12 > $O/racecheck_pair.txt; tail -4 $O/racecheck_pair.txt
timeout 300 python tools/comparators.py > $O/comparators.txt 2>&1; cat $O/comparators.txt
timeout 900 python bench.py --no-extras --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['attention'])"
