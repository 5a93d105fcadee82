#!/bin/bash
# compute-sanitizer memcheck over the sanitize target (smoke, group of 2 with the k_slab_* exchange, contacts + graph + adjacency)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/j_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/j_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/j_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/j_memcheck.log
