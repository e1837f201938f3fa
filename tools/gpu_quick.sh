#!/bin/bash
# quick GPU iteration: lift-related parity tests + short device-timed bench lines
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "${1:-lift or fused or full_size or golden or sort or mean or refine or pooling}" 2>&1 | tail -8
for args in "--variant 0" "--variant 4" "--variant 1" "--variant 2"; do
  timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu $args > gpurun_out/q.json 2>gpurun_out/q.err || tail -5 gpurun_out/q.err
  python tools/brief.py gpurun_out/q.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 64 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 10 --no-e2e --no-cpu > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather -s 12 -c 1 -o gpurun_out/lift_prof -f python bench.py --steps 4 --warmup 10 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
