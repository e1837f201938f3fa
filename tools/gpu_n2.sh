#!/bin/bash
# 2-GPU box: NCCL tests + the driver's N=2 command for both arms
set -u
mkdir -p gpurun_out
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "view_sharded or dist or nccl or push" 2>&1 | tail -3
n=2
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532"
( time timeout 900 $L bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/n2.json 2> gpurun_out/n2.err ) 2>&1 | grep real
tail -2 gpurun_out/n2.err
python -c "
import json; d=json.load(open('gpurun_out/n2.json')); print('ours N=2 value', d['value'], 'ms', d['ms_per_step'], 'host_us', d['host_us_per_step'], 'e2e', d['e2e']['value']); print({k: d['viewshard'][k] for k in ('ms_1gpu','ms_per_step','speedup_vs_1gpu')}); print('mask', d['mask_gemm']['us'], d['mask_gemm']['achieved'])"
( time timeout 600 $L bench.py --impl reference --gpus $n --steps 3 --warmup 1 > gpurun_out/n2_ref.json 2>> gpurun_out/n2.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/n2_ref.json')); print('ref N=2', d['value'], d['cpu_baseline']['cores'])"
