#!/bin/bash
set -u
tag=${1:-f2}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
[ $rc -eq 0 ] && python tools/sass_hashes.py node-cueball_b200/libcueball_b200.so --write gpurun_out/${tag}_validated_sass.json \
	"md5 of each kernel's SASS (tools/sass_hashes.py) of the build whose -m gpu suite passed on a B200 in this session ($tag)" > /dev/null
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload c2 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[c2] ms/step=%.4f e2e=%.3g (%.3f ms, pcie %.2f)' % (d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['frac_of_pcie']))"
bash tools/gpu_multi.sh 2 ${tag}_n2
