#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/tests.txt 2>&1; tail -8 $O/tests.txt
timeout 900 python bench.py --no-extras > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step','clocks')}); print(d['e2e']['value']); print(d['roofline']['other_share_of_step'], d['roofline']['attention'])
PY
