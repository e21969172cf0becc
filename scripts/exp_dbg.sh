#!/bin/bash
for v in 0 15; do MPGNN_TC_EXP=$v timeout 300 python scripts/exp_tc_dbg.py 2>&1 | tail -14; done
