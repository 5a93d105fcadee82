#!/bin/bash
# radix ranking with the four-instruction ballot step: A/B on the 16M workload against the C++ ballot loop
# (-DGSC_SORT_RANK_GENERIC), then the GPU suite (the sort tests cover 8..32 key bits) on the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/ab_variants.sh run rank4 rankgen rank4 2>&1 | tee gpurun_out/v_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -n 3 gpurun_out/v_pytest.log
timeout 300 python bench.py --config 2 --no-cpu-baseline > gpurun_out/v_bench_c2.json 2> gpurun_out/v_bench_c2.err; echo "config 2 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/v_bench_c2.json"))
print(2, round(d["ms_per_step"], 4), f'{d["value"]:.3e}', {k: v["ms"] for k, v in d["stages"].items()})
PY
