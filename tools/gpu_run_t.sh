#!/bin/bash
# final tree, 1 GPU: full GPU suite, the same against the checked library, smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log
tail -3 gpurun_out/t_pytest.log
GSC_LIB=$PWD/gunship_rs_b200/variants/lib_checked.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest_checked.log 2>&1; echo "checked pytest rc=$?" >> gpurun_out/t_pytest_checked.log
tail -3 gpurun_out/t_pytest_checked.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/t_smoke.log
