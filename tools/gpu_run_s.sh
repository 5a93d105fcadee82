#!/bin/bash
# round-2 GPU call S (8 GPUs): group / slab tests on 8 real GPUs, 16M scaling point at N=8, config 5 (64M spheres) at N=8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q -k "one_rank_per_gpu" > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
tail -4 gpurun_out/s_pytest.log
GSC_SLAB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --frames 2 > gpurun_out/s_bench8.json 2> gpurun_out/s_bench8.err; echo "bench8 rc=$?"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --config 5 --frames 1 --steps 20 --warmup 3 > gpurun_out/s_bench8_c5.json 2> gpurun_out/s_bench8_c5.err; echo "bench8 c5 rc=$?"
python - <<'PY'
import json
for f in ("s_bench8","s_bench8_c5"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        st=d.get("stages",{})
        print(f, round(d["ms_per_step"],4), f'{d["value"]:.3e}', 'e2e', f'{d["e2e"]["value"]:.3e}', round(d["e2e"]["ms_per_step"],3), {k:(v.get("ms_min_over_ranks"), v.get("ms_max_over_ranks")) for k,v in st.items()}, d["config"].get("frame0_pair_set"), d["config"].get("slab_calibration_busy_ms"), d["config"].get("colliders_per_rank"), d["config"].get("halo_colliders_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
grep "slab trace" gpurun_out/s_bench8.err | tail -2
tail -2 gpurun_out/s_bench8_c5.err
