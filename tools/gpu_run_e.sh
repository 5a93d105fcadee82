#!/bin/bash
# ncu: launch list + full capture of one step (same command exited 0 right before)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GSC_BENCH_GRAPH=0
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --frames 2"
$CMD > gpurun_out/e_plain.json 2> gpurun_out/e_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a.csv $CMD > gpurun_out/e_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/e_plain2.json 2> gpurun_out/e_plain2.err &&
ncu --set full --clock-control none --import-source on -s 21 -c 19 -o gpurun_out/prof_step_r2a $CMD > gpurun_out/e_ncu2.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/e_ncu2.log
