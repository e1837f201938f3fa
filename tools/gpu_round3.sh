#!/bin/bash
# mask GEMM (TMA / bf16x3) tests, timings and ncu captures; cfg5 sweep
set -u
mkdir -p gpurun_out
OUT=gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mask or layernorm" 2>&1 | tail -5
timeout 300 python tools/bench_mask.py > $OUT/r02_mask_gemm.jsonl 2> $OUT/mask.err; tail -3 $OUT/mask.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_mask_gemm.jsonl'):
    r=json.loads(l); print({k:(round(v['us'],1) if isinstance(v,dict) else v) for k,v in r.items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_logits_tma -s 2 -c 1 -o $OUT/r02_mask_tma -f python tools/exp_mask_tma.py 5000 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_logits_tma -s 2 -c 1 -o $OUT/r02_mask_tma_attn -f python tools/exp_mask_tma.py 5000 attn > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_logits_tma -s 2 -c 1 -o $OUT/r02_mask_tma_x3 -f python tools/exp_mask_tma.py 5000 split > /dev/null 2>&1
timeout 900 python tools/sweep.py > $OUT/r02_sweep.jsonl 2> $OUT/sweep.err; tail -3 $OUT/sweep.err; cat $OUT/r02_sweep.jsonl
ls -la $OUT | grep r02_mask
