#!/bin/bash
# what the driver does at round end, on one 8-GPU box: both arms at N GPUs
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "nearest or label_masks or view_sharded" 2>&1 | tail -5
for n in ${1:-8}; do
  if [ "$n" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"; fi
  ( time timeout 600 $L bench.py --impl reference --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_ref_$n.json 2> gpurun_out/scale_ref_$n.err ) 2>&1 | grep real
  python -c "import json; d=json.load(open('gpurun_out/scale_ref_$n.json')); print('ref N=$n', d['value'], d['cpu_baseline']['cores'])"
  ( time SD3D_STAGE_TIMES=1 timeout 900 $L bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err ) 2>&1 | grep real
  grep -a "stage ms" gpurun_out/scale_$n.err | sort | head -8
  python -c "
import json; d=json.load(open('gpurun_out/scale_$n.json')); print('ours N=$n value', d['value'], 'ms', d['ms_per_step'], 'host_us', d['host_us_per_step'], 'e2e', d['e2e']['value']); print(d.get('viewshard'))"
done
