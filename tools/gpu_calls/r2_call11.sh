#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.txt 2>&1; tail -3 $O/tests.txt
timeout 600 python tools/comparators.py > $O/comparators.txt 2>&1; grep attention $O/comparators.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; cat $O/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')}); print(d.get('attention'))"
