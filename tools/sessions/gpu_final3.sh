#!/bin/bash
# tools/gpu_final3.sh [tag] -- two GPUs: the GPU suite on GPU 0, then bench.py under torchrun on both.
set -u
tag=${1:-g1}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
[ $rc -ne 0 ] && grep -E "^(FAILED|ERROR)|Error" gpurun_out/${tag}_pytest.log | head -10
bash tools/gpu_multi.sh 2 ${tag}_n2
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_n2_bench.log').read().strip().splitlines()[-1])
print('by rank:', d.get('ms_per_step_by_rank'))
PY
