#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py tests/test_gpu_trainer.py -q -p no:cacheprovider -x > gpurun_out/r2h_pytest.log 2>&1
echo "pytest exit: $?"; tail -3 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-candidates --no-api --steps 16 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench exit: $?"; tail -2 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "hop frac", d["extra"]["hop_roofline"]["frac"])
for k, v in d["extra"]["kernels"].items():
    print("  %-28s %.3f ms  %.2f" % (k, v["ms_per_launch"], v.get("frac_of_bound", 0)))
PY
