#!/bin/bash
# round-2 GPU call G (2 GPUs): preflight of the calibrated multi-GPU bench (cfg 4) and cfg 5 at a reduced size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/g_bench2.json 2> gpurun_out/g_bench2.err; echo "bench2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --config 5 --colliders 4000000 --frames 1 --steps 20 --warmup 3 > gpurun_out/g_bench2_c5.json 2> gpurun_out/g_bench2_c5.err; echo "bench2 c5 rc=$?"
timeout 300 python bench.py --config 5 --colliders 4000000 --frames 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench1_c5.json 2> gpurun_out/g_bench1_c5.err; echo "bench1 c5 rc=$?"
python - <<'PY'
import json
for f in ("g_bench2","g_bench2_c5","g_bench1_c5"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        st=d.get("stages",{})
        print(f, round(d["ms_per_step"],4), f'{d["value"]:.3e}', 'e2e', f'{d["e2e"]["value"]:.3e}', {k:(v.get("ms") or (v.get("ms_min_over_ranks"), v.get("ms_max_over_ranks"))) for k,v in st.items()}, d["config"].get("frame0_pair_set"), d["config"].get("slab_calibration_busy_ms"), d["config"].get("colliders_per_rank"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/g_bench2_c5.err
