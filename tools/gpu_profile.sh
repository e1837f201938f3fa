#!/bin/bash
# round-end evidence: full tests, bench line, ncu launch list + full captures of the top kernels, mask GEMM timings
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err; python tools/brief.py $OUT/bench.json
timeout 300 python bench.py --streams 1 --no-e2e --no-cpu > $OUT/bench_1stream.json 2>> $OUT/bench.err; python tools/brief.py $OUT/bench_1stream.json
timeout 300 python tools/bench_mask.py > $OUT/mask_gemm.jsonl 2>> $OUT/bench.err; cat $OUT/mask_gemm.jsonl | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 70 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 12 --warmup 10 --streams 1 --no-overlap --no-e2e --no-cpu > /dev/null 2>&1
python tools/launch_summary.py $OUT/launches.csv | tee $OUT/launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 4 -c 1 -o $OUT/gather_prof -f \
   python bench.py --steps 4 --warmup 4 --streams 1 --no-overlap --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_kernel -s 4 -c 1 -o $OUT/project_prof -f \
   python bench.py --steps 4 --warmup 4 --streams 1 --no-overlap --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_logits_tc -s 2 -c 1 -o $OUT/mask_tc_prof -f \
   python -c "
import sys; sys.path.insert(0,'.')
import torch, segdino3d_b200 as sd
from segdino3d_b200.synth import make_decoder_operands
q,mf = make_decoder_operands(5000,5000,256); q,mf=q.cuda(),mf.cuda()
for _ in range(4): sd.mask_logits(q,mf,precision='bf16')
torch.cuda.synchronize()" > /dev/null 2>&1
ls -la $OUT | tail -20
