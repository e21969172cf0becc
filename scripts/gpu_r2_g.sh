#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_search_bags.py tests/test_gpu_cli.py -q -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2g_pytest.log; tail -5 gpurun_out/r2g_pytest.log
bash scripts/gpu_c5.sh 1
