#!/bin/bash
# quick GPU iteration: lift-related parity tests + short device-timed bench lines
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "${1:-lift or fused or full_size or golden or sort or mean}" 2>&1 | tail -8
for args in "--variant 0" "--variant 1" "--variant 0 --run 16"; do
  timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu $args > gpurun_out/q.json 2>gpurun_out/q.err || tail -5 gpurun_out/q.err
  python tools/brief.py gpurun_out/q.json
done
