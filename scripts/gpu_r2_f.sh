#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_device_bag_diag.py > gpurun_out/r2f_diag.log 2>&1; echo "diag exit $?"; grep -E "first|Error" gpurun_out/r2f_diag.log | cut -c1-400
timeout 900 python scripts/exp_c5_shape.py 100000 19 > gpurun_out/r2f_c5shape.log 2>&1; echo "c5 shape exit $?"; grep -vE "^test loss" gpurun_out/r2f_c5shape.log | cut -c1-260 | tail -40
