#!/bin/bash
set -u
tag=${1:-s5}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
[ $rc -eq 0 ] && python tools/sass_hashes.py node-cueball_b200/libcueball_b200.so --write gpurun_out/${tag}_validated_sass.json \
	"md5 of each kernel's SASS (tools/sass_hashes.py) of the build whose -m gpu suite passed on a B200 in this session ($tag)" > /dev/null
for w in c2 c3; do
	timeout 600 python bench.py --workload $w --no-extra > gpurun_out/${tag}_bench_$w.log 2> gpurun_out/${tag}_bench_$w.err
	echo "[bench $w] rc=$? :: $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_$w.log').read().strip().splitlines()[-1]); r=d['roofline']
    print('ms/step=%.4f host=%.4f scan=%.4f frac=%.3f step_frac=%.3f e2e=%.3g e2e_ms=%.3f pcie_frac=%.3f parity=%s' % (d['ms_per_step'], d['host_enqueue_ms_per_step'], r['kernel_ms']['scan'], r['frac'], r['step']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['frac_of_pcie'], d['parity']['ok']))
except Exception as e:
    print('unparsed', e)
PY
)"
done
CB_DEBUG_E2E=1 timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -25
t0=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 3 > gpurun_out/${tag}_bench_ref.log 2>&1
echo "[bench reference, 5 steps] rc=$? $(( $(date +%s) - t0 ))s :: $(tail -c 600 gpurun_out/${tag}_bench_ref.log)"
