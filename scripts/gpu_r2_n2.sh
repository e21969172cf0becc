#!/bin/bash
# 2-GPU pass: NCCL decision-equality test, tcgen05 + search tests on the new build, configs[4] search at N=2
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tcgen05.py tests/test_gpu_cli.py tests/test_gpu_search_bags.py -q -p no:cacheprovider > gpurun_out/r2n2_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2n2_pytest.log; tail -6 gpurun_out/r2n2_pytest.log
bash scripts/gpu_c5.sh 2
