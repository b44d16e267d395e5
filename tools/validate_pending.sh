#!/bin/bash
# tools/validate_pending.sh -- run on a B200 box (e.g. `gpurun --timeout 1500 -- 'bash tools/validate_pending.sh'`):
# everything that was written after round 1's GPU budget ran out, in the order NEXT.md gives.
# Each step prints one summary line; logs go to gpurun_out/pending_*.log.
set -u
mkdir -p gpurun_out
run() { # name, command...
	local name=$1; shift
	local t0=$(date +%s)
	"$@" > gpurun_out/pending_$name.log 2>&1
	local rc=$?
	echo "[$name] rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -1 gpurun_out/pending_$name.log | cut -c1-160)"
}
export CUDA_DEVICE_MAX_CONNECTIONS=32
run suite_default      timeout 600 python -m pytest tests -x -q -m gpu
run queue_and_fchain   env CB_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_queue_unvalidated.py tests/test_gpu_fchain_unvalidated.py -q
run suite_fchain       env CB_FCHAIN=1 timeout 600 python -m pytest tests -x -q -m gpu
run suite_fused        env CB_FUSED=1 timeout 600 python -m pytest tests -x -q -m gpu
run suite_fused_fchain env CB_FUSED=1 CB_FCHAIN=1 timeout 600 python -m pytest tests -x -q -m gpu
run suite_graph        env CB_GRAPH=1 timeout 600 python -m pytest tests -x -q -m gpu
run suite_graph_all    env CB_GRAPH=1 CB_FUSED=1 CB_FCHAIN=1 timeout 600 python -m pytest tests -x -q -m gpu
for v in "" "CB_FCHAIN=1" "CB_FUSED=1" "CB_FUSED=1 CB_FCHAIN=1" "CB_GRAPH=1" "CB_GRAPH=1 CB_FUSED=1 CB_FCHAIN=1"; do
	for w in c2 c3; do
		env $v python bench.py --workload $w --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('[bench %-24s %s] value=%.3g ms/step=%.4f host=%.4f launches/step=%.1f e2e=%.3g' % ('$v' or 'default', d['config']['workload'], d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['gpu_launches']/d['steps'], d['e2e']['value']))"
	done
done
for v in "" "CB_FUSED=1 CB_FCHAIN=1"; do
	echo "--- timeline c2 ($v)"; env $v timeout 120 python tools/pipeline_timeline.py c2 5 60 2>&1 | tail -6
done
