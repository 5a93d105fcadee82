#!/bin/bash
# k_cell_table_permute with 1 / 2 / 4 sorted positions per thread (record gathers in flight together): A/B at 16M
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/ab_variants.sh run t1 t2 t2c6 t4 2>&1 | tee gpurun_out/w_ab.txt
python - <<'PY'
import json
for n in ("t1", "t2", "t2c6", "t4"):
    d = json.load(open(f"gpurun_out/ab_{n}.json"))
    print(n, d["config"]["frame0_pair_set"], d["config"]["assign_order"])
PY
