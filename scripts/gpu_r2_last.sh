#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/last_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/last_pytest.log; tail -3 gpurun_out/last_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/last_smoke.log 2>&1; echo "smoke exit: $?"
timeout 1500 python bench.py > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
echo "bench exit: $?"; tail -2 gpurun_out/last_bench.err; python -c "
import json
d=json.load(open('gpurun_out/last_bench.json'))
print(d['ms_per_step'], d['value'], d['clocks'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])
"
