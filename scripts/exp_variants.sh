#!/bin/bash
# builds variants of proj_tcgen05.cu by sed-editing constants into a temp copy, timing each
SRC=mpgnn-metapath-graph-neural-network_b200/csrc/proj_tcgen05.cu
cp $SRC /tmp/proj_orig.cu
run() { python mpgnn-metapath-graph-neural-network_b200/_build.py > /dev/null 2>&1 && EXP_TAG="$1" timeout 300 python scripts/exp_tc.py 2>&1 | tail -1; }
run "base(prefetch4)"
sed -i 's/constexpr int kPrefetch = 4;/constexpr int kPrefetch = 2;/' $SRC; run "prefetch2"
cp /tmp/proj_orig.cu $SRC
sed -i 's/constexpr int kPrefetch = 4;/constexpr int kPrefetch = 6;/' $SRC; run "prefetch6"
cp /tmp/proj_orig.cu $SRC
python mpgnn-metapath-graph-neural-network_b200/_build.py > /dev/null 2>&1
