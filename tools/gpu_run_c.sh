#!/bin/bash
# round-2 GPU call C (2 GPUs): all GPU tests (group API, device-driven exchange on 2 ranks), 2-GPU bench, pairs A/B on 1 GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c_bench2.json 2> gpurun_out/c_bench2.err; echo "bench2 rc=$?"
GSC_EXCHANGE=host timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c_bench2_host.json 2> gpurun_out/c_bench2_host.err; echo "bench2 host rc=$?"
for v in default nofuse tiled; do
  case $v in default) E="";; nofuse) E="GSC_FUSE_SS=0";; tiled) E="GSC_PAIRS_TILED=1";; esac
  env $E timeout 300 python bench.py --steps 10 --warmup 3 --frames 2 --no-cpu-baseline > gpurun_out/c_ab_$v.json 2> gpurun_out/c_ab_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for f in ("c_bench2","c_bench2_host","c_ab_default","c_ab_nofuse","c_ab_tiled"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        st=d.get("stages",{})
        print(f, round(d["ms_per_step"],4), f'{d["value"]:.3e}', 'e2e', f'{d["e2e"]["value"]:.3e}', {k:(v.get("ms") or v.get("ms_max_over_ranks")) for k,v in st.items()}, d["config"].get("frame0_pair_set"))
    except Exception as e:
        print(f, "ERR", e)
PY
