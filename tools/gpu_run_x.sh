#!/bin/bash
# final tree, 1 GPU: GPU suite, smoke(), the default bench line, ncu launch list + full capture of one step
# (plain launches), then the suite against the checked library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log
tail -n 3 gpurun_out/x_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/x_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?"
export GSC_BENCH_GRAPH=0
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --frames 2"
$CMD > gpurun_out/x_plain.json 2> gpurun_out/x_plain.err &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2e.csv $CMD > gpurun_out/x_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -s 80 -c 19 -o gpurun_out/prof_step_r2e $CMD > gpurun_out/x_ncu2.log 2>&1
echo "full rc=$?"; tail -n 2 gpurun_out/x_ncu2.log
unset GSC_BENCH_GRAPH
left=$((660 - SECONDS)); [ $left -lt 60 ] && left=60
GSC_LIB=$PWD/gunship_rs_b200/variants/lib_checked.so timeout $left python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest_checked.log 2>&1; echo "checked pytest rc=$?" >> gpurun_out/x_pytest_checked.log
tail -n 3 gpurun_out/x_pytest_checked.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/x_bench.json"))
print(json.dumps({k:d[k] for k in ("value","ms_per_step","p50_step_ms","roofline","cpu_baseline","config1_bench_scene","stages","clocks","gpu_launches")})[:2600])
print(d["e2e"])
print(d["config"])
PY
