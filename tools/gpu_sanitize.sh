#!/bin/bash
# tools/gpu_sanitize.sh [tag] -- compute-sanitizer racecheck + synccheck + memcheck over the golden-vector GPU
# tests (small inputs: the tools slow kernels down 10-100x).  Logs go to gpurun_out/<tag>_sanitize_*.txt;
# the summaries are kept under profiles/.
set -u
tag=${1:-san}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
SEL="tests/test_gpu_parity.py::test_golden_vectors tests/test_gpu_parity.py::test_scalar_overloaded_and_empty tests/test_gpu_parity.py::test_sequential_fallback_matches_oracle tests/test_gpu_parity.py::test_plan_rebalance_reference_kat tests/test_gpu_parity.py::test_slot_table_golden tests/test_gpu_queue.py::test_device_queue_on_golden_scripts"
python -m pytest $SEL --collect-only -q 2>/dev/null | tail -3
for tool in racecheck synccheck memcheck; do
	t0=$(date +%s)
	timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
		python -m pytest $SEL -x -q > gpurun_out/${tag}_sanitize_$tool.txt 2>&1
	echo "[$tool] rc=$? $(( $(date +%s) - t0 ))s :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_sanitize_$tool.txt | tail -1) :: $(grep -E 'passed|failed' gpurun_out/${tag}_sanitize_$tool.txt | tail -1)"
done
