#!/bin/bash
# 8-GPU diagnostic of the view-sharded path (cfg4): 1-GPU time, then 8 GPUs with the p2p exchange + per-stage times
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests -m gpu -q --tb=short -x -k "one_call or view_sharded" 2>&1 | tail -4
( time timeout 600 python bench.py --mode viewshard --workload cfg4 --steps 10 --warmup 3 > gpurun_out/vs1_cfg4.json 2> gpurun_out/vs1.err ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/vs1_cfg4.json')); print('1 GPU', d['ms_per_step'], d['config'])"
for ex in p2p; do
  ( time SD3D_STAGE_TIMES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 8 --mode viewshard --workload cfg4 --exchange $ex --steps 20 --warmup 3 > gpurun_out/vs8_$ex.json 2> gpurun_out/vs8_$ex.err ) 2>&1 | grep real
  grep "stage ms" gpurun_out/vs8_$ex.err | sort | head -8
  python -c "import json; d=json.load(open('gpurun_out/vs8_$ex.json')); print('8 GPU $ex', d['ms_per_step'])"
done
