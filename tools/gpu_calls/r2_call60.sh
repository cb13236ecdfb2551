#!/bin/bash
# final library (after the pool_partial change): full GPU suite, smoke, bench N=1 with all legs
O=gpurun_out/r2at; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/gpu_tests.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee -a $O/gpu_tests.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'], d['roofline']['other_share_of_step'])
P
