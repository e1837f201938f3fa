#!/bin/bash
# Runs on the B200 box under gpurun: parity tests, a short bench, and the ncu launch list.
# Everything lands in gpurun_out/ (merged back into the repo checkout).
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
echo "== pytest (all but tensor-core mask kernel)" 
timeout 1200 python -m pytest tests -m gpu -q --tb=short -k "not tcgen05 and not forward_head" > $OUT/pytest_main.log 2>&1
echo "exit $?" >> $OUT/pytest_main.log; tail -15 $OUT/pytest_main.log
echo "== pytest (tcgen05 mask kernel, separate process)"
timeout 600 python -m pytest tests -m gpu -q --tb=short -k "tcgen05 or forward_head" > $OUT/pytest_tc.log 2>&1
echo "exit $?" >> $OUT/pytest_tc.log; tail -15 $OUT/pytest_tc.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 200 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
for v in 1; do
  timeout 300 python bench.py --steps 200 --warmup 10 --variant $v --no-e2e --no-cpu > $OUT/bench_v$v.json 2>> $OUT/bench.err; python tools/brief.py $OUT/bench_v$v.json
done
for r in 8 16 64; do
  timeout 300 python bench.py --steps 200 --warmup 10 --run $r --no-e2e --no-cpu > $OUT/bench_run$r.json 2>> $OUT/bench.err; python tools/brief.py $OUT/bench_run$r.json
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 12 --warmup 10 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
tail -2 $OUT/ncu_bench.log
echo "== ncu full capture of the lift kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 12 -c 2 -o $OUT/lift_prof -f \
   python bench.py --steps 4 --warmup 10 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ls -la $OUT
