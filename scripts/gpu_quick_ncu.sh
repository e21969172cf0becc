#!/bin/bash
# quick loop + one ncu --set full capture of the tensor-core kernels (fwd, wgrad, dgrad of one hop)
bash scripts/gpu_quick.sh "$@" || exit 1
TAG=${NCU_TAG:-quick}
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-candidates $*"
ncu --set full --clock-control none --import-source on -k regex:'gemm_rows_tc|wgrad_tc' -s 9 -c 3 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit: $?"
