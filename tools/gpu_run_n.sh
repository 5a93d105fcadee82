#!/bin/bash
# round-2 GPU call N (1 GPU): final bench line + ncu launch list + full capture of one step (slot order, plain launches)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"
export GSC_BENCH_GRAPH=0
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --frames 2"
$CMD > gpurun_out/n_plain.json 2> gpurun_out/n_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2c.csv $CMD > gpurun_out/n_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/n_plain2.json 2> gpurun_out/n_plain2.err &&
ncu --set full --clock-control none --import-source on -s 80 -c 19 -o gpurun_out/prof_step_r2c $CMD > gpurun_out/n_ncu2.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/n_ncu2.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/n_bench.json"))
print(json.dumps({k:d[k] for k in ("value","ms_per_step","p50_step_ms","roofline","cpu_baseline","config1_bench_scene","stages")})[:2500])
print(d["e2e"])
print(d["config"])
PY
