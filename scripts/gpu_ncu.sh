#!/bin/bash
# ncu captures for profiles/: launch list (durations) + one --set full capture of a kernel.
# usage: gpu_ncu.sh <tag> <kernel-regex> [bench args...]
TAG=$1; KREGEX=$2; shift 2
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-candidates $*"
$CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list exit: $?"
$CMD > gpurun_out/ncu_plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 3 -c 3 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit: $?"
ls -la gpurun_out | tail -12
