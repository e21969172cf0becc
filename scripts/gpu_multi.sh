#!/bin/bash
# Multi-GPU check (run with gpurun --gpus N): search fan-out over NCCL + the bench at N ranks.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit: $?"; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | cut -c1-700
