#!/bin/bash
# round-2 ncu captures of the C4 hop (compact form): launch list + one --set full capture of the 9 kernels of ONE hop.
# usage (under gpurun): bash scripts/gpu_ncu_r2.sh <tag> [bench args...]
TAG=${1:-r02}; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-candidates $*"
$CMD > gpurun_out/ncu_plain_${TAG}.json 2> gpurun_out/ncu_plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_${TAG}.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list exit: $?"
# hop kernels only; skip the three warm-up hops (9 matching launches each), capture the next hop
ncu --set full --clock-control none --import-source on -k regex:'gemm_rows_tc|wgrad_tc_kernel|spmm_gather|spmm_accumulate|gather_gated' \
    -s 27 -c 9 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit: $?"
ls -la gpurun_out | grep ${TAG}
