#!/bin/bash
mkdir -p gpurun_out
for W in 296 592 1184; do
MPGNN_WAVE_CTAS=$W timeout 600 python - <<'PY' 2>> gpurun_out/r2k.err
import json, os, sys, torch
sys.path.insert(0, ".")
import bench
out = bench.candidate_scoring(0, 1, torch.device("cuda", 0), None)
print(os.environ["MPGNN_WAVE_CTAS"], out["candidates_per_s"], out["seconds"], out["seconds_one_candidate_alone"])
PY
done
