#!/bin/bash
# compute-sanitizer over the smoke test and the GPU parity tests that exercise this round's kernel changes
# (accumulator count loops, in-lane run check, lane records, work order, fixed first streams)
set -u
mkdir -p gpurun_out
T=${1:-r02s}
for tool in memcheck racecheck; do
  echo "== $tool smoke"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_${tool}_smoke.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/${T}_sanitizer_${tool}_smoke.log | head -5
done
echo "== memcheck gpu tests"
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path or span_by_span or segment_by_segment or foreign or golden or composition or truncated" > gpurun_out/${T}_sanitizer_memcheck_gpu_tests.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_memcheck_gpu_tests.log | head -5
echo "== racecheck gpu tests"
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path or span_by_span or foreign" > gpurun_out/${T}_sanitizer_racecheck_gpu_tests.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_racecheck_gpu_tests.log | head -5
