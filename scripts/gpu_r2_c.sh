#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py tests/test_gpu_trainer.py -q -p no:cacheprovider -s > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c_pytest.log; grep -E "passed|failed|FAILED|tf32x3 \{|fp32 \{|999|C2 trace" gpurun_out/r2c_pytest.log | tail -40
timeout 900 python bench.py --no-e2e --no-cpu-baseline --no-candidates --steps 8 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench exit: $?"; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "hop frac", d["extra"]["hop_roofline"]["frac"])
for k, v in d["extra"]["kernels"].items():
    print("  %-28s %.3f ms x %d" % (k, v["ms_per_launch"], v["calls"]))
PY
