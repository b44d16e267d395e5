#!/bin/bash
# tools/gpu_check.sh [tag] [quick] -- one GPU session: the -m gpu suite, device-resident bench lines for c3 and
# c2, one `ncu --set full` capture of the scan kernel at C3, scan-grid knobs.  Logs and reports go to gpurun_out/.
set -u
tag=${1:-chk}
quick=${2:-}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
# the device code that just ran: pinned by tests/test_validated_sass.py once copied to profiles/
[ $rc -eq 0 ] && python tools/sass_hashes.py node-cueball_b200/libcueball_b200.so --write gpurun_out/${tag}_validated_sass.json \
	"md5 of each kernel's SASS (tools/sass_hashes.py) of the build whose -m gpu suite passed on a B200 in this session ($tag)" > /dev/null
show() { # file
	python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; k=r['kernel_ms']
    print('value=%.4g ms/step=%.4f scan=%.4f (isolated %.4f) frac=%.3f step_frac=%.3f resolve=%.4f emit=%.4f host=%.4f e2e=%.3g pcie_frac=%s parity=%s' % (
        d['value'], d['ms_per_step'], k['scan'], k.get('scan_isolated', 0), r['frac'], r['step']['frac'], k['resolve'], k['emit'],
        d['host_enqueue_ms_per_step'], d['e2e']['value'], d['e2e'].get('frac_of_pcie'), d.get('parity', {}).get('ok')))
except Exception as e:
    print('unparsed', e)
PY
}
for w in c3 c2; do
	timeout 900 python bench.py --workload $w --no-extra > gpurun_out/${tag}_bench_$w.log 2> gpurun_out/${tag}_bench_$w.err
	echo "[bench $w] rc=$? :: $(show gpurun_out/${tag}_bench_$w.log)"
done
[ -n "$quick" ] && exit 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -c 4 -f -o gpurun_out/${tag}_scan_c3 \
	python bench.py --workload c3 --no-extra --steps 3 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "[ncu] rc=$?"
# scan-grid knobs at C3 (measurement only)
for v in "CB_SCAN_TPW=2" "CB_SCAN_TPW=3" "CB_SCAN_TPW=64"; do
	env $v timeout 300 python bench.py --workload c3 --no-extra > gpurun_out/${tag}_knob.log 2>/dev/null
	echo "[knob $v c3] $(show gpurun_out/${tag}_knob.log)"
done
# compile-time geometry (rebuilt on the box, default restored at the end)
for x in "-DCB_SCAN_WARPS=6 -DCB_SCAN_NST=4" "-DCB_SCAN_WARPS=12 -DCB_SCAN_NST=2" "-DCB_SCAN_WARPS=4 -DCB_SCAN_NST=3 -DCB_SCAN_MINBLOCKS=4"; do
	make -s -C node-cueball_b200/csrc -B EXTRA="$x" > /dev/null 2>&1 || { echo "[build $x] failed"; continue; }
	for v in "CB_SCAN_TPW=1" "CB_SCAN_TPW=2"; do
		env $v timeout 300 python bench.py --workload c3 --no-extra > gpurun_out/${tag}_knob.log 2>/dev/null
		echo "[build $x | $v c3] $(show gpurun_out/${tag}_knob.log)"
	done
done
make -s -C node-cueball_b200/csrc -B > /dev/null 2>&1
