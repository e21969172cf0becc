#!/bin/bash
# GPU tests, then the ncu launch list + full capture of one kernel family.
# usage: gpu_tests_ncu.sh <tag> <kernel-regex> [bench args...]
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
bash scripts/gpu_ncu.sh "$@"
