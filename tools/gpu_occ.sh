#!/bin/bash
set -u
mkdir -p gpurun_out
cd /root/repo
for args in "--variant 1 --streams 2" "--variant 65 --streams 2" "--variant 65 --streams 3" "--variant 1 --streams 3"; do
  timeout 300 python bench.py --steps 400 --warmup 10 --no-e2e --no-cpu --no-viewshard --no-mask $args > gpurun_out/q.json 2>gpurun_out/q.err || tail -5 gpurun_out/q.err
  echo "$args"; python tools/brief.py gpurun_out/q.json
done
