#!/bin/bash
# quick iteration loop on the GPU box: tcgen05 + parity tests, then the kernel-only bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py -x -q -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?"; tail -15 gpurun_out/pytest_quick.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-candidates "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit: $?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("ms/step", d["ms_per_step"], "value", d["value"])
for k, v in d["extra"]["kernels"].items():
    print("  %-24s %.3f ms" % (k, v["ms_per_launch"]))
PY
