#!/bin/bash
# after the wide-epilogue forward: all GPU tests, smoke, default bench, ncu captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/final_pytest.log; tail -4 gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke exit: $?"; tail -1 gpurun_out/final_smoke.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench exit: $?"; tail -2 gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json
bash scripts/gpu_ncu_r2.sh r02
