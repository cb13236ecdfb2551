#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/tests.txt 2>&1; tail -6 $O/tests.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step','clocks','preheat')}); print(d['e2e']['value']); print(d.get('c3')); print(d.get('c5'))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 400 $O/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench_n2.json').read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['e2e']['value']); print(d.get('c3')); print(d.get('c5'))
PY
