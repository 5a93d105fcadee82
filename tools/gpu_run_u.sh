#!/bin/bash
# 256-bit record loads / stores: A/B on the 16M workload (default build, -DGSC_NO_LDST256, 4 resident narrow CTAs),
# then the GPU suite on the default build and the config 2 / config 3 bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/ab_variants.sh run d256 no256 n4 d256 no256 2>&1 | tee gpurun_out/u_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log
tail -n 3 gpurun_out/u_pytest.log
for c in 2 3; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/u_bench_c$c.json 2> gpurun_out/u_bench_c$c.err; echo "config $c rc=$?"
done
python - <<'PY'
import json
for c in (2, 3):
    try:
        d = json.load(open(f"gpurun_out/u_bench_c{c}.json"))
        print(c, round(d["ms_per_step"], 4), f'{d["value"]:.3e}', "e2e", f'{d["e2e"]["value"]:.3e}', {k: v["ms"] for k, v in d["stages"].items()})
    except Exception as e:
        print(c, "failed", e)
PY
