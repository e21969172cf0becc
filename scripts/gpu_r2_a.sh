#!/bin/bash
# round 2, first GPU pass: all GPU tests, smoke, bag-parity diagnostic
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -x --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2a_pytest.log; tail -40 gpurun_out/r2a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke exit: $?"; tail -2 gpurun_out/r2a_smoke.log
timeout 600 python scripts/exp_bag_parity.py > gpurun_out/r2a_bag_parity.log 2>&1; echo "bag parity exit: $?"; tail -40 gpurun_out/r2a_bag_parity.log
