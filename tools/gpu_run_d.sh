#!/bin/bash
# round-2 GPU call D (1 GPU): new tests, bench (default / slot order / graph at 1M), config-1 C++ bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py tests/test_gpu_parity.py tests/test_lifecycle.py tests/test_contacts.py -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -8 gpurun_out/d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --frames 2 --no-cpu-baseline --slot-order > gpurun_out/d_bench_slot.json 2> gpurun_out/d_bench_slot.err; echo "slot rc=$?"
timeout 300 python bench.py --config 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_c2.json 2> gpurun_out/d_bench_c2.err; echo "c2 rc=$?"
g++ -std=c++17 -O2 -I include -I host benches/grid_collision.cpp -L gunship_rs_b200 -lgunship_collision -Wl,-rpath,$PWD/gunship_rs_b200 -o /tmp/grid_collision_bench && /tmp/grid_collision_bench 1000 500 > gpurun_out/d_cpp_bench.txt 2>&1
cat gpurun_out/d_cpp_bench.txt
python - <<'PY'
import json
for f in ("d_bench","d_bench_slot","d_bench_c2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        st=d.get("stages",{})
        print(f, round(d["ms_per_step"],4), f'{d["value"]:.3e}', 'e2e', f'{d["e2e"]["value"]:.3e}', d["e2e"].get("with_host_marshal"), {k:v.get("ms") for k,v in st.items()}, d["config"].get("frame0_pair_set"), d.get("config1_bench_scene"))
    except Exception as e:
        print(f, "ERR", e)
PY
