#!/bin/bash
set -u
tag=${1:-d2}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
[ $rc -eq 0 ] && python tools/sass_hashes.py node-cueball_b200/libcueball_b200.so --write gpurun_out/${tag}_validated_sass.json \
	"md5 of each kernel's SASS (tools/sass_hashes.py) of the build whose -m gpu suite passed on a B200 in this session ($tag)" > /dev/null
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step=%.4f by_rank=%s host=%.4f' % (d['ms_per_step'], d.get('ms_per_step_by_rank'), d['host_enqueue_ms_per_step']))
except Exception as e:
    print('unparsed', e)
PY
}
for dp in 8 10 12; do
	env CB_PIPELINE_DEPTH=$dp CB_BENCH_FAST=1 CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-extra > gpurun_out/mx.log 2>/dev/null
	echo "[N=1 c3 depth=$dp] $(show gpurun_out/mx.log)"
	env CB_PIPELINE_DEPTH=$dp CB_BENCH_FAST=1 CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload c2 --no-extra > gpurun_out/mx.log 2>/dev/null
	echo "[N=1 c2 depth=$dp] $(show gpurun_out/mx.log)"
done
for dp in 8 12; do
	env CB_PIPELINE_DEPTH=$dp CB_BENCH_FAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 \
		bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/mx.log 2>/dev/null
	echo "[N=2 c3 depth=$dp peers] $(show gpurun_out/mx.log)"
done
