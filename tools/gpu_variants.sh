#!/bin/bash
# bench the lift-kernel experiment variants (device-timed only); each arg is a quoted bench.py option string
set -u
mkdir -p gpurun_out
for v in "$@"; do
  timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu $v > gpurun_out/q.json 2>gpurun_out/q.err || tail -5 gpurun_out/q.err
  echo -n "[$v] "; python tools/brief.py gpurun_out/q.json
done
