#!/bin/bash
# configs[4] full search through bench.py at N GPUs: bash scripts/gpu_c5.sh N
N=${1:-1}
mkdir -p gpurun_out
ARGS="--gpus $N --search-only"
if [ "$N" == "1" ]; then
  timeout 1500 python bench.py $ARGS > gpurun_out/c5_n$N.json 2> gpurun_out/c5_n$N.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py $ARGS > gpurun_out/c5_n$N.json 2> gpurun_out/c5_n$N.err
fi
echo "c5 N=$N exit: $?"; grep -vE "^test loss|^W[0-9]|^\*\*\*" gpurun_out/c5_n$N.err | tail -5
python - <<PY
import json
d = json.load(open("gpurun_out/c5_n$N.json"))
c = d["extra"]["search_c5"]
print({k: v for k, v in c.items() if k not in ("log", "final_dict")})
print("seconds", d["value"])
PY
