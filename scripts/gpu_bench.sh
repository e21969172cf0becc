#!/bin/bash
# Runs on the GPU box: parity tests, then the bench (tiny plumbing check, then the C4 line).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload tiny --steps 8 --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
echo "tiny exit: $?"; tail -3 gpurun_out/bench_tiny.err; cat gpurun_out/bench_tiny.json
timeout 1500 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "c4 exit: $?"; tail -5 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
