#!/bin/bash
# compute-sanitizer over the smoke test (fast path, general kernel, span / segment passes, stored, PNG rows)
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
