#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py tests/test_gpu_trainer.py -q -p no:cacheprovider -x > gpurun_out/r2m_pytest.log 2>&1
echo "pytest exit: $?"; tail -3 gpurun_out/r2m_pytest.log
timeout 600 python - <<'PY' 2> gpurun_out/r2m.err
import json, sys, torch
sys.path.insert(0, ".")
import bench
out = bench.candidate_scoring(0, 1, torch.device("cuda", 0), None)
print({k: v for k, v in out.items() if k in ("candidates_per_s", "seconds", "seconds_one_candidate_alone", "f1_planted_metapath")})
PY
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-candidates --no-api --steps 16 2> gpurun_out/r2m_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', d['ms_per_step'])
for k,v in d['extra']['kernels'].items(): print('  %-28s %.3f'%(k,v['ms_per_launch']))
"
