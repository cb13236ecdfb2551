#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
for shape in "1,1,417,4.0" "1,2,1568,1.0" "1,2,480,1.0"; do
echo "=== synccheck shape $shape"
timeout 300 compute-sanitizer --tool synccheck --print-limit 200 python -c "
import sys; sys.path.insert(0,'.')
from tests import kernel_checks as kc
import torch
B,H,S,pk=[float(x) for x in '$shape'.split(',')]
kc.check_attention(int(B),int(H),int(S),peaky=pk,seed=4); torch.cuda.synchronize(); print('ok', flush=True)
" 2>&1 | grep -v "Host Frame\|=========         in\|=========     Saved\|Device Frame\|^========= $" | awk '{c[$0]++} END {for (k in c) print c[k], k}' | sort -rn | head -12
done | tee $O/synccheck_shapes.txt
