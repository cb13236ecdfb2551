#!/bin/bash
# end-of-round evidence with the final library: GPU suite, smoke, bench N=1 (all legs), reference arm, other configs,
# ncu launch list of the bench command
O=gpurun_out/r2ar; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_tests.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee -a $O/gpu_tests.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python tools/bench_configs.py > $O/other_configs.jsonl 2> $O/other.err; echo "configs rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python - <<P
import json
d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])
print('c3', d.get('c3')); print('cpu', d.get('cpu_baseline'))
r=json.loads(open('$O/bench_reference_arm.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
P
cat $O/other_configs.jsonl | cut -c1-200
