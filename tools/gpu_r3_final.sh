#!/bin/bash
# third session of round 2: A/B of the last two builds, the evidence run (tests, bench, reference arm, launch list, ncu
# captures), then racecheck over the K4 tests
set -u
mkdir -p gpurun_out
T=${1:-r03f}
bash tools/gpu_ab_list.sh ${T} ab/t2.so ab/t3.so
bash tools/gpu_r2_final.sh ${T}
echo "== racecheck gpu tests"
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path or span_by_span or foreign" > gpurun_out/${T}_sanitizer_racecheck_gpu_tests.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_racecheck_gpu_tests.log | head -5
