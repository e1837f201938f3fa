#!/bin/bash
# round-2 evidence: bench lines, ncu launch list + full captures of the top kernels, mask GEMM timings, op timings
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 900 python bench.py > $OUT/r02_bench_cfg2.json 2> $OUT/bench.err; tail -2 $OUT/bench.err; python tools/brief.py $OUT/r02_bench_cfg2.json
timeout 600 python bench.py --workload cfg3 --no-viewshard > $OUT/r02_bench_cfg3.json 2>> $OUT/bench.err; python tools/brief.py $OUT/r02_bench_cfg3.json
timeout 300 python bench.py --variant 0 --no-e2e --no-cpu --no-viewshard > $OUT/r02_bench_cfg2_exact.json 2>> $OUT/bench.err; python tools/brief.py $OUT/r02_bench_cfg2_exact.json
timeout 300 python bench.py --variant 32769 --no-e2e --no-cpu --no-viewshard > $OUT/r02_bench_cfg2_staged.json 2>> $OUT/bench.err; python tools/brief.py $OUT/r02_bench_cfg2_staged.json
timeout 300 python bench.py --streams 1 --no-e2e --no-cpu --no-viewshard > $OUT/r02_bench_cfg2_1stream.json 2>> $OUT/bench.err; python tools/brief.py $OUT/r02_bench_cfg2_1stream.json
timeout 300 python tools/bench_mask.py > $OUT/r02_mask_gemm.jsonl 2>> $OUT/bench.err
timeout 300 python tools/bench_ops.py > $OUT/r02_ops.jsonl 2>> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 72 --csv --log-file $OUT/r02_launches.csv \
   python bench.py --steps 12 --warmup 10 --streams 1 --no-e2e --no-cpu --no-viewshard > /dev/null 2>&1
python tools/launch_summary.py $OUT/r02_launches.csv | tee $OUT/r02_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 6 -c 1 -o $OUT/r02_gather -f \
   python bench.py --steps 4 --warmup 4 --streams 1 --no-e2e --no-cpu --no-viewshard > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gather_staged|project_stage" -s 6 -c 2 -o $OUT/r02_gather_staged -f \
   python bench.py --steps 4 --warmup 4 --streams 1 --variant 32769 --no-e2e --no-cpu --no-viewshard > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 6 -c 1 -o $OUT/r02_gather_fp16 -f \
   python bench.py --workload cfg3 --steps 4 --warmup 4 --streams 1 --no-e2e --no-cpu --no-viewshard > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_logits_tc -s 2 -c 1 -o $OUT/r02_mask_tc -f \
   python -c "
import sys; sys.path.insert(0,'.')
import torch, segdino3d_b200 as sd
from segdino3d_b200.synth import make_decoder_operands
q,mf = make_decoder_operands(5000,5000,256); q,mf=q.cuda(),mf.cuda()
for _ in range(4): sd.mask_logits(q,mf,precision='bf16')
torch.cuda.synchronize()" > /dev/null 2>&1
ls -la $OUT | grep r02
