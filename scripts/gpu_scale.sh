#!/bin/bash
# The driver's scaling launch for N ranks with the default bench flags (e2e included).
N=${1:-8}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_scale_n$N.json 2> gpurun_out/bench_scale_n$N.err
echo "bench N=$N exit: $?"; tail -3 gpurun_out/bench_scale_n$N.err; python - <<PY
import json
d = json.load(open("gpurun_out/bench_scale_n$N.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], "cand", d["extra"]["candidate_scoring"]["candidates_per_s"])
PY
