#!/bin/bash
# round-2 GPU call F (2 GPUs): preflight of the multi-GPU bench paths (cfg 4 and cfg 5 small), 1-GPU bench after the pairs fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_slabs.py -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err; echo "bench1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err; echo "bench2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --config 5 --n 4000000 --frames 1 --steps 20 --warmup 3 > gpurun_out/f_bench2_c5.json 2> gpurun_out/f_bench2_c5.err; echo "bench2 c5 rc=$?"
timeout 300 python bench.py --config 5 --n 4000000 --frames 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench1_c5.json 2> gpurun_out/f_bench1_c5.err; echo "bench1 c5 rc=$?"
timeout 300 python bench.py --config 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ("f_bench1","f_bench2","f_bench2_c5","f_bench1_c5","f_bench_c2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        st=d.get("stages",{})
        print(f, round(d["ms_per_step"],4), f'{d["value"]:.3e}', 'e2e', f'{d["e2e"]["value"]:.3e}', d["config"].get("assign_order"), {k:(v.get("ms") or v.get("ms_max_over_ranks")) for k,v in st.items()}, d["config"].get("frame0_pair_set"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/f_bench2_c5.err
