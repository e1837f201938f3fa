#!/bin/bash
# final mask GEMM evidence: tests, timings, ncu captures of the three variants
set -u
mkdir -p gpurun_out
OUT=gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_goldens.py -x -q -m gpu -k "mask or layernorm or out_norm or head" 2>&1 | tail -3
timeout 300 python tools/bench_mask.py > $OUT/r02_mask_gemm.jsonl 2> $OUT/mask.err; tail -3 $OUT/mask.err
for m in "" attn split; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_logits_tma -s 2 -c 1 -o $OUT/r02_mask_tma${m:+_$m} -f python tools/exp_mask_tma.py 5000 $m > /dev/null 2>&1
done
ls -la $OUT | grep r02_mask
