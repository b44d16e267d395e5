#!/bin/bash
# tools/gpu_multi.sh N [tag] -- bench.py under torchrun on N GPUs exactly as the driver launches it, then the
# reference arm's shape check (rank 0 only prints).
set -u
N=${1:-2}
tag=${2:-m$N}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
	bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
echo "[bench N=$N] rc=$? $(( $(date +%s) - t0 ))s"
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1])
    r=d['roofline']
    print('main: value=%.4g ms/step=%.4f per-GPU=%.4g scan=%.4f frac=%.3f step_frac=%.3f e2e=%.3g parity=%s launches=%d' % (d['value'], d['ms_per_step'], d['value']/d['n_gpus'], r['kernel_ms']['scan'], r['frac'], r['step']['frac'], d['e2e']['value'], d['parity'], d['gpu_launches']))
    print('global_counts:', d.get('global_counts'))
    print('nccl:', json.dumps(d.get('nccl_allreduce'))[:600])
    print('one_queue:', json.dumps(d.get('one_queue'))[:700])
    for k in ('c4','c5'):
        x=d.get(k); print(k, 'value=%.4g ms/step=%.4f parity=%s gc=%s' % (x['value'], x['ms_per_step'], x['parity']['ok'], x['global_counts']['equals_sum_of_ranks']))
except Exception as e:
    print('unparsed', e)
PY
