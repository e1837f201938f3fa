#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "cfg3 or cfg4 or batched or nearest or label" 2>&1 | tail -5
timeout 300 python tools/exp_staged.py fp16 0 32768 1 2>&1 | tail -4
timeout 300 python tools/bench_mask.py > gpurun_out/mask_gemm.jsonl 2>gpurun_out/mg.err; tail -2 gpurun_out/mg.err
python - <<'PY'
import json
for l in open('gpurun_out/mask_gemm.jsonl'):
    d=json.loads(l); print({k:(round(v['us'],1) if isinstance(v,dict) else v) for k,v in d.items() if k!='flop'})
PY
timeout 600 python bench.py --workload cfg3 --no-viewshard --no-cpu > gpurun_out/b_cfg3.json 2> gpurun_out/b_cfg3.err; tail -2 gpurun_out/b_cfg3.err; python tools/brief.py gpurun_out/b_cfg3.json
