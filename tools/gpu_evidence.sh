#!/bin/bash
# end-of-round evidence on one GPU: full suite, smoke, bench lines (cfg2 default, cfg3), reference arm, operator / mask timings,
# ncu launch list of the step and full captures of the TMA mask GEMM
set -u
mkdir -p gpurun_out
OUT=gpurun_out
cd /root/repo
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | tail -6 | tee $OUT/r02_gputests_1gpu_box.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/r02_bench_cfg2.json 2> $OUT/bench.err; tail -2 $OUT/bench.err; python tools/brief.py $OUT/r02_bench_cfg2.json
timeout 600 python bench.py --workload cfg3 --no-viewshard --no-mask > $OUT/r02_bench_cfg3.json 2>> $OUT/bench.err; python tools/brief.py $OUT/r02_bench_cfg3.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference_arm.json 2>> $OUT/bench.err; cut -c1-200 $OUT/r02_bench_reference_arm.json
timeout 300 python tools/bench_mask.py > $OUT/r02_mask_gemm.jsonl 2>> $OUT/bench.err
timeout 300 python tools/bench_ops.py > $OUT/r02_ops.jsonl 2>> $OUT/bench.err
for m in "" attn split; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_logits_tma -s 2 -c 1 -o $OUT/r02_mask_tma${m:+_$m} -f python tools/exp_mask_tma.py 5000 $m > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 72 --csv --log-file $OUT/r02_launches.csv \
   python bench.py --steps 12 --warmup 10 --streams 1 --no-e2e --no-cpu --no-viewshard --no-mask > /dev/null 2>&1
python tools/launch_summary.py $OUT/r02_launches.csv | tee $OUT/r02_launches_summary.txt | tail -12
ls -la $OUT | grep r02_ | wc -l
