#!/bin/bash
# tools/gpu_final.sh [tag] -- what the driver runs at round end, in one session: the -m gpu suite (+ SASS hashes of
# the build that passed), smoke(), the default bench line with its extra keys.
set -u
tag=${1:-fin}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
[ $rc -eq 0 ] && python tools/sass_hashes.py node-cueball_b200/libcueball_b200.so --write gpurun_out/${tag}_validated_sass.json \
	"md5 of each kernel's SASS (tools/sass_hashes.py) of the build whose -m gpu suite passed on a B200 in this session ($tag)" > /dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$(date +%s)
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.log 2> gpurun_out/${tag}_bench_n1.err
echo "[bench default] rc=$? $(( $(date +%s) - t0 ))s :: $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_n1.log').read().strip().splitlines()[-1])
    r=d['roofline']; c2=d['c2']
    print('c3 value=%.4g ms/step=%.4f scan=%.4f frac=%.3f step_frac=%.3f e2e=%.3g pcie_frac=%.3f parity=%s | c2 ms/step=%.4f host=%.4f e2e=%.3g (%.3f ms, pcie %.2f) parity=%s step_frac=%.3f | stress=%.3g exact=%s | cpu=%.3g' % (
        d['value'], d['ms_per_step'], r['kernel_ms']['scan'], r['frac'], r['step']['frac'], d['e2e']['value'], d['e2e']['frac_of_pcie'], d['parity']['ok'],
        c2['ms_per_step'], c2['host_enqueue_ms_per_step'], c2['e2e']['value'], c2['e2e']['ms_per_step'], c2['e2e']['frac_of_pcie'], c2['parity']['ok'], c2['roofline']['step']['frac'],
        d['nonmonotone_stress']['value'], d['nonmonotone_stress']['bit_exact_vs_oracle'], d['cpu_baseline']['value']))
except Exception as e:
    print('unparsed', e)
PY
)"
