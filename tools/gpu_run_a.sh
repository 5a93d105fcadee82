#!/bin/bash
# round-2 GPU call A: parity (incl. full-size), bench, legacy A/B, CUB yardstick
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
GSC_PAIRS_LEGACY=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_legacy.json 2> gpurun_out/a_bench_legacy.err; echo "legacy rc=$?"
timeout 120 tools/cub_sort_yardstick 16777216 27 20 > gpurun_out/a_cub27.json 2>&1
timeout 120 tools/cub_sort_yardstick 16777216 32 20 > gpurun_out/a_cub32.json 2>&1
timeout 120 tools/cub_sort_yardstick 2097152 24 20 > gpurun_out/a_cub24_2m.json 2>&1
cat gpurun_out/a_cub27.json
python - <<'PY'
import json
for f in ("a_bench","a_bench_legacy"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["ms_per_step"], d["value"], {k:v["ms"] for k,v in d["stages"].items()}, d["config"].get("frame0_pair_set"))
    except Exception as e:
        print(f, "ERR", e)
PY
