#!/bin/bash
O=gpurun_out/r2ag; mkdir -p $O
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "two_ranks_nccl" 2>&1 | tail -2 | tee $O/nccl_test.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err ) 2>&1 | grep real
python -c "
import json
d=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value']); print('c3', d.get('c3')); print('c5', str(d.get('c5'))[:600])"
( time timeout 900 python bench.py > $O/bench_n1_full.json 2> $O/bench_n1_full.err ) 2>&1 | grep real
tail -c 1500 $O/bench_n1_full.json
( time timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err ) 2>&1 | grep real
tail -c 700 $O/bench_ref.json
