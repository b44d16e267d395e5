#!/bin/bash
# N=2: is the main pass of the full line slower than the timed pass alone (CB_BENCH_FAST)?  3 samples each.
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step=%.4f by_rank=%s' % (d['ms_per_step'], d.get('ms_per_step_by_rank')))
except Exception as e:
    print('unparsed', e)
PY
}
run2() {
	env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 \
		bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/mx.log 2>/dev/null
	echo "[N=2 $*] $(show gpurun_out/mx.log)"
}
for i in 1 2 3; do run2 CB_BENCH_FAST=1; run2 CB_BENCH_FAST=0; done
run2 CB_BENCH_FAST=0 CB_BENCH_SAMPLER=0
run2 CB_BENCH_FAST=0 CB_BENCH_SAMPLER=0
