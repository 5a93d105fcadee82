#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --frames 2"
$CMD > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_pairs_tiled -s 1 -c 1 -o gpurun_out/prof_pairs_tiled_v1 $CMD > gpurun_out/b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/b_ncu.log
