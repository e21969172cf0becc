#!/bin/bash
# device search pipeline tests + RGCN baseline tests + the C5 search at N=1 (timed alone)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_search_bags.py tests/test_gpu_search.py tests/test_gpu_rgcn_baseline.py tests/test_gpu_cli.py -q -p no:cacheprovider -s > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2e_pytest.log; grep -E "passed|failed|FAILED|Error|search [0-9.]+ s" gpurun_out/r2e_pytest.log | tail -30
timeout 2400 python - > gpurun_out/r2e_c5.json 2> gpurun_out/r2e_c5.err <<'PY'
import json, sys, time, torch
sys.path.insert(0, ".")
import bench
t0 = time.time()
out = bench.search_c5(0, 1, torch.device("cuda", 0), None)
out["wall_s"] = time.time() - t0
print(json.dumps(out))
PY
echo "c5 exit: $?"; tail -5 gpurun_out/r2e_c5.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_c5.json"))
print({k: v for k, v in d.items() if k not in ("log", "final_dict")})
print("\n".join(d["log"][:40]))
PY
