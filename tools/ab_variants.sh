#!/bin/bash
# Development helper: A/B kernel variants of libgunship_collision.so in ONE gpurun call.
#
#   here (no GPU):   tools/ab_variants.sh build  p24:-DGSC_PAIR_MIN_CTAS=24  priv16:-DGSC_PAIR_PRIV=16 ...
#   on the GPU box:  gpurun -- 'bash tools/ab_variants.sh run p24 priv16'
#
# `build` compiles one .so per "name:flags" into gunship_rs_b200/variants/ (git-ignored, but it travels with the
# gpurun snapshot); `run` benches each on the 16M workload (GSC_LIB selects the library, see _native.py) and prints
# the per-stage CUDA-event times.  Compile-time knobs: GSC_PAIR_MIN_CTAS, GSC_PAIR_THREADS, GSC_PAIR_PRIV,
# GSC_NARROW_CTAS (gsc_kernels.cuh).  Remove the variants directory before committing a measurement run's result.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
V="$ROOT/gunship_rs_b200/variants"
mode="$1"; shift
if [ "$mode" = build ]; then
    mkdir -p "$V"
    for spec in "$@"; do
        name="${spec%%:*}"; flags="${spec#*:}"
        ( cd "$ROOT/gunship_rs_b200/csrc" && /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
            -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC $flags -Xptxas -v -shared \
            -o "$V/lib_$name.so" gsc_api.cu -lcudart 2> "$V/build_$name.log" ) &
    done
    wait
    for spec in "$@"; do name="${spec%%:*}"; echo "$name: $(grep -A3 '7k_pairs' "$V/build_$name.log" | grep Used)"; done
elif [ "$mode" = run ]; then
    mkdir -p "$ROOT/gpurun_out"
    for name in "$@"; do
        GSC_LIB="$V/lib_$name.so" python "$ROOT/bench.py" --no-cpu-baseline --steps 8 --frames 2 > "$ROOT/gpurun_out/ab_$name.json" 2> "$ROOT/gpurun_out/ab_$name.err"
        python - "$ROOT/gpurun_out/ab_$name.json" "$name" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], round(d["device_ms_per_step"], 4), {k: v["ms"] for k, v in d["stages"].items()})
PY
    done
else
    echo "usage: $0 build name:flags... | run name..."; exit 2
fi
