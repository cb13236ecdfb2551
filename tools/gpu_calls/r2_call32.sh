#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
timeout 600 compute-sanitizer --tool synccheck --print-limit 3 python tools/sanitize_small.py attention > $O/synccheck_full.txt 2>&1
grep -v "Host Frame\|=========         in\|^=========     Saved" $O/synccheck_full.txt | head -40
