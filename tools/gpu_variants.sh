#!/bin/bash
# bench the lift-kernel experiment variants (device-timed only)
set -u
mkdir -p gpurun_out
for v in "$@"; do
  timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu --variant $v > gpurun_out/q.json 2>gpurun_out/q.err || tail -5 gpurun_out/q.err
  python tools/brief.py gpurun_out/q.json
done
