#!/bin/bash
set -u
mkdir -p gpurun_out
for args in "--variant 0 --streams 2" "--variant 1 --streams 2" "--variant 1 --streams 3" "--variant 1 --streams 4" "--variant 33 --streams 2"; do
  timeout 300 python bench.py --steps 400 --warmup 10 --no-e2e --no-cpu --no-viewshard $args > gpurun_out/q.json 2>gpurun_out/q.err || tail -5 gpurun_out/q.err
  echo "$args"; python tools/brief.py gpurun_out/q.json
done
timeout 300 python bench.py --workload cfg3 --steps 400 --warmup 10 --no-e2e --no-cpu --no-viewshard --variant 1 > gpurun_out/q.json 2>gpurun_out/q.err || tail -5 gpurun_out/q.err
echo "cfg3 variant 1"; python tools/brief.py gpurun_out/q.json
