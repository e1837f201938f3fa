#!/bin/bash
# what the driver does at round end: full GPU suite, smoke, default bench line (+ reference arm)
set -u
mkdir -p gpurun_out
cd /root/repo
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err; python tools/brief.py gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2>> gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_ref.json
