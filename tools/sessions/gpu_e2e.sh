#!/bin/bash
# tools/gpu_e2e.sh [tag] -- the host-buffer entry point at 1M claims: the GPU suite first (host code changed),
# then a per-chunk timeline (CB_DEBUG_E2E) and the chunk-plan knobs side by side.
set -u
tag=${1:-e1}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
[ $rc -eq 0 ] && python tools/sass_hashes.py node-cueball_b200/libcueball_b200.so --write gpurun_out/${tag}_validated_sass.json \
	"md5 of each kernel's SASS (tools/sass_hashes.py) of the build whose -m gpu suite passed on a B200 in this session ($tag)" > /dev/null
[ $rc -ne 0 ] && grep -E "^(FAILED|ERROR)|Error" gpurun_out/${tag}_pytest.log | head -10
run() { # label, env...
	label=$1; shift
	echo "== $label"
	env "$@" CUDA_VISIBLE_DEVICES=0 timeout 120 python tools/e2e_breakdown.py 1048576 0 2>gpurun_out/${tag}_$label.err | grep -v "^bound" | cut -c1-400
}
run default X=0
env CB_DEBUG_E2E=1 CUDA_VISIBLE_DEVICES=0 timeout 120 python tools/e2e_breakdown.py 1048576 0 2>&1 | grep "cb e2e" | tail -12
run split CB_H2D_SPLIT=1
run pct50 CB_CHUNK_PCT=50
run pct30 CB_CHUNK_PCT=30
run floor32k CB_CHUNK_FLOOR=32768
run floor128k CB_CHUNK_FLOOR=131072
run pct50_floor128k CB_CHUNK_PCT=50 CB_CHUNK_FLOOR=131072
run split_pct50 CB_H2D_SPLIT=1 CB_CHUNK_PCT=50
run chunk256k CB_CHUNK=262144
run lanes12 CB_PIPELINE_DEPTH=12
echo "== 16M"
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/e2e_breakdown.py 16777216 0 2>/dev/null | grep -v "^bound\|rotating" | cut -c1-300
CB_H2D_SPLIT=1 CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/e2e_breakdown.py 16777216 0 2>/dev/null | grep "admit_batch" | cut -c1-300
