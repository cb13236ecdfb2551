#!/bin/bash
# bench line of the final tree at N GPUs (N from the environment; default 1)
N=${N:-1}
O=gpurun_out/r2ay; mkdir -p $O
if [ $N = 1 ]; then
  timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
  timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference_arm.json 2> $O/bench_ref.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
fi
echo "rc=$?"
python - <<P
import json
d=json.loads(open('$O/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks'])
print('c3', d.get('c3')); print('cpu', d.get('cpu_baseline'))
P
