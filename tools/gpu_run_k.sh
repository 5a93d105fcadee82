#!/bin/bash
# round-2 GPU call K (1 GPU): full GPU suite on the final tree; the same suite against the -DGSC_CHECKED library (internal
# bounds checks: the substitute for compute-sanitizer, which is closed on this pool), also with the tiled pairs kernel;
# config 5 (64M spheres) on ONE GPU for the fingerprint the 8-GPU run must reproduce
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -4 gpurun_out/k_pytest.log
GSC_LIB=$PWD/gunship_rs_b200/variants/lib_checked.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest_checked.log 2>&1; echo "checked pytest rc=$?" >> gpurun_out/k_pytest_checked.log
tail -4 gpurun_out/k_pytest_checked.log
GSC_PAIRS_TILED=1 GSC_FUSE_SS=1 GSC_LIB=$PWD/gunship_rs_b200/variants/lib_checked.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/k_pytest_checked_tiled.log 2>&1; echo "checked tiled pytest rc=$?" >> gpurun_out/k_pytest_checked_tiled.log
tail -4 gpurun_out/k_pytest_checked_tiled.log
timeout 900 python bench.py --config 5 --frames 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench1_c5.json 2> gpurun_out/k_bench1_c5.err; echo "bench c5 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/k_bench1_c5.json"))
print(round(d["ms_per_step"],3), f'{d["value"]:.3e}', d["e2e"]["value"], d["config"].get("frame0_pair_set"), d["config"].get("assign_order"), {k:v["ms"] for k,v in d["stages"].items()})
PY
