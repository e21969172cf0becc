#!/bin/bash
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2d_pytest.log; tail -8 gpurun_out/r2d_pytest.log
timeout 1500 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench exit: $?"; tail -3 gpurun_out/r2d_bench.err; cut -c1-1500 gpurun_out/r2d_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_ref.json 2> gpurun_out/r2d_ref.err
echo "ref exit: $?"; tail -3 gpurun_out/r2d_ref.err; cat gpurun_out/r2d_ref.json
bash scripts/gpu_ncu_r2.sh r02a
