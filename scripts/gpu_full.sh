#!/bin/bash
# Full round check on the GPU box: all GPU tests, smoke, bench (default), reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?"; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit: $?"; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "ref exit: $?"; cat gpurun_out/bench_reference.json
