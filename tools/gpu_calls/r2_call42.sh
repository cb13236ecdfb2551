#!/bin/bash
O=gpurun_out/r2ae; mkdir -p $O
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee $O/gpu_tests.txt
grep -q failed $O/gpu_tests.txt && exit 1
timeout 300 python tools/profile_dapt.py 100 > $O/dapt.txt 2>&1; head -14 $O/dapt.txt
timeout 300 python tools/profile_step.py --depth 0 --batch 1 --iters 5 --table > $O/batch1_table.txt 2>&1; cat $O/batch1_table.txt | tail -22
timeout 900 python bench.py --no-extras --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['attention'], d.get('clocks'))"
