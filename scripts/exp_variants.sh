#!/bin/bash
# Pipeline-stage ablation of the tcgen05 projection kernel.  Needs a library built with
#   MPGNN_NVCC_EXTRA=-DMPGNN_TC_EXPERIMENT python -c "import __graft_entry__ as g; g.build()"
# MPGNN_TC_EXP bits: 1 epilogue off (ld + release only), 2 converter math/LDS off, 4 only the hi*hi MMA,
# 8 no TMA loads, 16 no TMA stores.  Results are WRONG by construction; only the times matter.
for v in ${EXP_LIST:-0 1 2 4 8 16 3 11 7 15}; do
  MPGNN_TC_EXP=$v EXP_TAG="exp=$v" timeout 300 python scripts/exp_tc.py 2>&1 | tail -1
done
