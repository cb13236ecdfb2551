#!/bin/bash
O=gpurun_out/r2aj; mkdir -p $O
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee $O/gpu_tests.txt
grep -q failed $O/gpu_tests.txt && exit 1
timeout 600 python tools/bench_configs.py > $O/other_configs.jsonl 2> $O/other_configs.err; cut -c1-200 $O/other_configs.jsonl
timeout 600 python tools/bench_configs.py --siblings > $O/sibling_configs.jsonl 2>> $O/other_configs.err; cut -c1-200 $O/sibling_configs.jsonl
timeout 300 python tools/comparators.py > $O/comparators.txt 2>&1; cat $O/comparators.txt | tail -14
timeout 300 python tools/profile_dapt.py 100 > $O/dapt.txt 2>&1
timeout 300 python -c "
import sys; sys.path.insert(0,'.')
from simple_tad_b200 import efficiency
efficiency.main()" > $O/efficiency_batch1.txt 2>&1; tail -8 $O/efficiency_batch1.txt
