#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "view_sharded or pool_superpoints_keeps" 2>&1 | tail -4
n=2
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $n --steps 20 --warmup 3 --no-e2e > gpurun_out/n2.json 2> gpurun_out/n2.err ) 2>&1 | grep real
tail -3 gpurun_out/n2.err
python -c "
import json; d=json.load(open('gpurun_out/n2.json')); print('ours N=2 value', d['value'], 'ms', d['ms_per_step'], 'host_us', d['host_us_per_step']); print(d.get('viewshard'))"
