#!/bin/bash
set -u
mkdir -p gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mask or layernorm" 2>&1 | tail -5
timeout 300 python tools/bench_mask.py > gpurun_out/r02_mask_gemm.jsonl 2> gpurun_out/mask.err; tail -3 gpurun_out/mask.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_mask_gemm.jsonl'):
    r=json.loads(l); print({k:(round(v['us'],1) if isinstance(v,dict) else v) for k,v in r.items()})
PY
