#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/tests.txt 2>&1; tail -4 $O/tests.txt
timeout 900 python bench.py --no-extras > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 200 $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['attention'])"
timeout 900 python tools/bench_configs.py --iters 10 > $O/configs.jsonl 2> $O/configs.err; cut -c1-260 $O/configs.jsonl
timeout 600 python tools/bench_configs.py --siblings --iters 10 > $O/siblings.jsonl 2>> $O/configs.err; cut -c1-200 $O/siblings.jsonl
timeout 300 python tools/profile_dapt.py > $O/profile_dapt.txt 2>&1; tail -25 $O/profile_dapt.txt
timeout 300 python tools/comparators.py > $O/comparators.txt 2>&1; cat $O/comparators.txt
