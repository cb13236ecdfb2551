#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py attention > $O/memcheck_att.txt 2>&1
grep -v "^=========     Host Frame\|^=========         in\|^=========$" $O/memcheck_att.txt | head -60
