#!/bin/bash
# Experiment driver: tcgen05 tests, then the hop bench per weight-gradient variant (MPGNN_WGRAD_EXP=1: register-staged).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcgen05.py -q -p no:cacheprovider -x 2>&1 | tail -4
for mode in ${MODES:-0 1}; do
  MPGNN_WGRAD_EXP=$mode timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-candidates --no-api --steps 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['extra']['kernels']
print('mode $mode ms/step', round(d['ms_per_step'],3), 'wgrad', round(k['wgrad_tn_tcgen05']['ms_per_launch'],3), round(k['wgrad_tn_compact_tcgen05']['ms_per_launch'],3), 'dgrad', round(k['dgrad_nt_tcgen05']['ms_per_launch'],3))
"
done
