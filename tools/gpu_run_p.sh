#!/bin/bash
# round-2 GPU call P (4 GPUs): multi-process device-exchange tests at 2 and 4 ranks (2M colliders), 16M scaling point at N=4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 3 --frames 2 > gpurun_out/p_bench4.json 2> gpurun_out/p_bench4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/p_bench4.json"))
st=d.get("stages",{})
print(round(d["ms_per_step"],4), f'{d["value"]:.3e}', 'e2e', f'{d["e2e"]["value"]:.3e}', {k:(v.get("ms_min_over_ranks"), v.get("ms_max_over_ranks")) for k,v in st.items()}, d["config"].get("frame0_pair_set"), d["config"].get("colliders_per_rank"))
PY
