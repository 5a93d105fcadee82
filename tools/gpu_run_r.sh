#!/bin/bash
# round-2 GPU call R (2 GPUs): halo push dealt out round-robin + final tunings: multi-GPU tests, N=2 bench, 1-GPU bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
tail -3 gpurun_out/r_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --frames 2 > gpurun_out/r_bench2.json 2> gpurun_out/r_bench2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --frames 2 --no-cpu-baseline > gpurun_out/r_bench1.json 2> gpurun_out/r_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for f in ("r_bench2","r_bench1"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    st=d.get("stages",{})
    print(f, round(d["ms_per_step"],4), f'{d["value"]:.3e}', 'e2e', f'{d["e2e"]["value"]:.3e}', {k:(v.get("ms") or (v.get("ms_min_over_ranks"), v.get("ms_max_over_ranks"))) for k,v in st.items()}, d["config"].get("frame0_pair_set"))
PY
