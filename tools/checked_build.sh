#!/bin/bash
# Builds gunship_rs_b200/variants/lib_checked.so: the library with its internal bounds checks compiled in (-DGSC_CHECKED,
# see GSC_CHECK in gsc_kernels.cuh).  compute-sanitizer is closed on the development pool; this is the substitute:
#   tools/checked_build.sh && GSC_LIB=$PWD/gunship_rs_b200/variants/lib_checked.so python -m pytest tests -m gpu -q
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$ROOT/gunship_rs_b200/variants"
cd "$ROOT/gunship_rs_b200/csrc"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true \
    -ftz=false -Xcompiler -fPIC -DGSC_CHECKED -shared -o ../variants/lib_checked.so gsc_api.cu -lcudart
echo built "$ROOT/gunship_rs_b200/variants/lib_checked.so"
