#!/bin/bash
# 4-GPU box: view-sharded cfg4 with balanced / equal-count view ranges (same box), NCCL p2p test
set -u
mkdir -p gpurun_out
cd /root/repo
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
for b in 1 0; do
  SD3D_VIEW_BALANCE=$b SD3D_STAGE_TIMES=1 timeout 600 $L bench.py --mode viewshard --workload cfg4 --exchange p2p --gpus 4 --steps 20 --warmup 3 > gpurun_out/vs4_$b.json 2> gpurun_out/vs4_$b.err
  echo "balance=$b"; python -c "
import json; d=json.load(open('gpurun_out/vs4_$b.json')); print(d['value'], d['ms_per_step'])"; grep -a "stage ms" gpurun_out/vs4_$b.err | sort | head -4
done
timeout 300 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -2
