#!/bin/bash
# N=2 diagnosis: which ingredient costs the pipelined step its overlap?  (timed pass only)
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step=%.4f host=%.4f by_rank=%s' % (d['ms_per_step'], d['host_enqueue_ms_per_step'], d.get('ms_per_step_by_rank')))
except Exception as e:
    print('unparsed', e)
PY
}
run2() { # env...
	env "$@" CB_BENCH_FAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 \
		bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/mx.log 2>/dev/null
	echo "[N=2 $*] $(show gpurun_out/mx.log)"
}
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_pipeline_streams.py -x -q -m gpu 2>&1 | tail -2
run2 CB_BENCH_PEERS=0
run2 CB_BENCH_PEERS=1
run2 CB_BENCH_PEERS=0
run2 CB_BENCH_PEERS=1
run2 CB_BENCH_PEERS=1 CB_PEERS_DEBUG=4
