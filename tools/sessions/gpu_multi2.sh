#!/bin/bash
# what costs the N>1 main pass its overlap?  the same C3 pass: N=1 plain, N=1 posting to its own board, N=2 with and without boards
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_pipeline_streams.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step=%.4f per-GPU=%.4g host=%.4f launches=%d by_rank=%s gc=%s' % (d['ms_per_step'], d['value']/d['n_gpus'], d['host_enqueue_ms_per_step'], d['gpu_launches'], d.get('ms_per_step_by_rank'), (d.get('global_counts') or {}).get('equals_sum_of_ranks')))
except Exception as e:
    print('unparsed', e)
PY
}
for v in "CB_BENCH_PEERS=0" "CB_BENCH_PEERS=self"; do
	env $v CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-extra > gpurun_out/mx.log 2>/dev/null
	echo "[N=1 $v] $(show gpurun_out/mx.log)"
done
for v in "CB_BENCH_PEERS=0" "CB_BENCH_PEERS=1"; do
	env $v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
		bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/mx.log 2>/dev/null
	echo "[N=2 $v] $(show gpurun_out/mx.log)"
done
