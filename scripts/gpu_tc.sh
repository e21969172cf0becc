#!/bin/bash
# tcgen05 path only (guarded by timeouts: a hang must not take the box down)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit: $?" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
if grep -q "passed" gpurun_out/pytest_tc.log && ! grep -q "failed" gpurun_out/pytest_tc.log; then
  timeout 900 python bench.py --precision tf32x3 --no-cpu-baseline > gpurun_out/bench_c4_tf32x3.json 2> gpurun_out/bench_c4_tf32x3.err
  echo "bench exit: $?"; tail -3 gpurun_out/bench_c4_tf32x3.err; cat gpurun_out/bench_c4_tf32x3.json
fi
