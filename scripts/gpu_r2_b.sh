#!/bin/bash
# all GPU tests (no -x), smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=12 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2b_pytest.log; tail -60 gpurun_out/r2b_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke exit: $?"; tail -2 gpurun_out/r2b_smoke.log
