#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python bench.py --no-extras --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['attention'], d.get('clocks'))"
