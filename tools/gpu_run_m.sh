#!/bin/bash
# round-2 GPU call M (1 GPU): light pairs tail (exact AABB test moved into k_narrow): full suite, checked suite, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -4 gpurun_out/m_pytest.log
GSC_LIB=$PWD/gunship_rs_b200/variants/lib_checked.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_group.py tests/test_contacts.py -m gpu -x -q > gpurun_out/m_pytest_checked.log 2>&1; echo "checked pytest rc=$?" >> gpurun_out/m_pytest_checked.log
tail -3 gpurun_out/m_pytest_checked.log
GSC_PAIRS_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/m_pytest_tiled.log 2>&1; echo "tiled pytest rc=$?" >> gpurun_out/m_pytest_tiled.log
tail -3 gpurun_out/m_pytest_tiled.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
GSC_PAIRS_TILED=1 timeout 600 python bench.py --steps 10 --warmup 3 --frames 2 --no-cpu-baseline > gpurun_out/m_bench_tiled.json 2> gpurun_out/m_bench_tiled.err; echo "bench tiled rc=$?"
python - <<'PY'
import json
for f in ("m_bench","m_bench_tiled"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["ms_per_step"],4), f'{d["value"]:.3e}', d["config"].get("assign_order"), {k:v["ms"] for k,v in d["stages"].items()}, d["config"].get("frame0_pair_set"))
    except Exception as e: print(f, "ERR", e)
PY
