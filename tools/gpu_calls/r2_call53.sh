#!/bin/bash
# 8-GPU bench line (headline + config-3 strong-scaling leg + config-5 sweep per rank) and the 2-rank NCCL test
O=gpurun_out/r2am; mkdir -p $O
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "rc=$?"; tail -3 $O/bench_n$N.err
python - <<P
import json
d=json.loads(open('$O/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
print(d.get('c3')); print(d.get('c5'))
P
