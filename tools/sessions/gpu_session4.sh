#!/bin/bash
# one GPU session: full -m gpu suite (+ SASS hashes), the default bench line with its extra keys, C2 knobs,
# launch list, ncu captures of the scan kernel at c2/c3/c4 sizes, compute-sanitizer.
set -u
tag=${1:-s4}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
echo "[pytest] rc=$rc :: $(tail -1 gpurun_out/${tag}_pytest.log)"
[ $rc -eq 0 ] && python tools/sass_hashes.py node-cueball_b200/libcueball_b200.so --write gpurun_out/${tag}_validated_sass.json \
	"md5 of each kernel's SASS (tools/sass_hashes.py) of the build whose -m gpu suite passed on a B200 in this session ($tag)" > /dev/null
t0=$(date +%s)
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_default.log 2> gpurun_out/${tag}_bench_default.err
echo "[bench default] rc=$? $(( $(date +%s) - t0 ))s :: $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_default.log').read().strip().splitlines()[-1])
    r=d['roofline']; c2=d['c2']
    print('c3 value=%.4g ms/step=%.4f scan=%.4f frac=%.3f step_frac=%.3f e2e=%.3g pcie_frac=%.3f parity=%s | c2 ms/step=%.4f host=%.4f e2e=%.3g parity=%s step_frac=%.3f | stress=%s | cpu=%.3g' % (
        d['value'], d['ms_per_step'], r['kernel_ms']['scan'], r['frac'], r['step']['frac'], d['e2e']['value'], d['e2e']['frac_of_pcie'], d['parity']['ok'],
        c2['ms_per_step'], c2['host_enqueue_ms_per_step'], c2['e2e']['value'], c2['parity']['ok'], c2['roofline']['step']['frac'], d['nonmonotone_stress'], d['cpu_baseline']['value']))
except Exception as e:
    print('unparsed', e)
PY
)"
t0=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 3 > gpurun_out/${tag}_bench_ref.log 2>&1
echo "[bench reference, 5 steps] rc=$? $(( $(date +%s) - t0 ))s :: $(tail -c 400 gpurun_out/${tag}_bench_ref.log)"
for v in "CB_PIPELINE_DEPTH=8" "CB_PIPELINE_DEPTH=8 CB_GRAPH=0" "CB_GRAPH=0" "CB_PIPELINE_DEPTH=3"; do
	env $v timeout 300 python bench.py --workload c2 --no-extra > gpurun_out/${tag}_knob.log 2>/dev/null
	echo "[c2 $v] $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_knob.log').read().strip().splitlines()[-1])
    print('ms/step=%.4f host=%.4f step_frac=%.3f e2e=%.3g' % (d['ms_per_step'], d['host_enqueue_ms_per_step'], d['roofline']['step']['frac'], d['e2e']['value']))
except Exception as e:
    print('unparsed', e)
PY
)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_c3.csv \
	python bench.py --no-extra --steps 3 --warmup 3 > gpurun_out/${tag}_launches.log 2>&1
echo "[launch list] rc=$? $(wc -l < gpurun_out/${tag}_launches_c3.csv) lines"
for w in c3 c2 c4; do
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|emit_kernel|resolve_kernel" -c 9 -f -o gpurun_out/${tag}_kernels_$w \
		python bench.py --workload $w --no-extra --steps 3 --warmup 3 > gpurun_out/${tag}_ncu_$w.log 2>&1
	echo "[ncu $w] rc=$?"
done
bash tools/gpu_sanitize.sh $tag
