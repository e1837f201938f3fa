#!/bin/bash
# one 8-GPU box: full GPU suite (the NCCL tests run), the driver's N=8 bench command, the configs[4] sweep on 8 GPUs
set -u
mkdir -p gpurun_out
cd /root/repo
nvidia-smi -L | wc -l
( time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | tail -6 | tee gpurun_out/r02_gputests_8gpu_box.txt
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
( time timeout 900 $L bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/scale8_20.json 2> gpurun_out/scale8_20.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/scale8_20.json')); print('N=8 steps 20: value', d['value'], 'ms', d['ms_per_step'], 'host_us', d['host_us_per_step'], 'e2e', d['e2e']['value']); print(d.get('viewshard'))"
( time timeout 900 $L bench.py --gpus 8 --no-viewshard > gpurun_out/scale8_default.json 2> gpurun_out/scale8_default.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/scale8_default.json')); print('N=8 default steps: value', d['value'], 'ms', d['ms_per_step'], 'host_us', d['host_us_per_step'])"
( time timeout 900 $L tools/sweep.py > gpurun_out/r02_sweep_8gpu.jsonl 2> gpurun_out/sweep8.err ) 2>&1 | grep real
tail -2 gpurun_out/sweep8.err; cat gpurun_out/r02_sweep_8gpu.jsonl
