#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python - <<'PY' 2> gpurun_out/r2j.err
import json, sys, torch
sys.path.insert(0, ".")
import bench
out = bench.candidate_scoring(0, 1, torch.device("cuda", 0), None)
print({k: v for k, v in out.items() if k != "config"})
PY
