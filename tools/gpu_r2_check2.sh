#!/bin/bash
# Round-2 check #2 on the GPU box: GPU tests (streaming, device sets, slot layouts), sanitizer over the new kernels.
set -u
mkdir -p gpurun_out
T=${1:-r02c}
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee gpurun_out/${T}_pytest.txt
echo "== memcheck (streaming + multi + layouts)"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_streaming.py tests/test_multi_device.py tests/test_slot_layouts.py -m gpu -x -q -k "not linear" 2>&1 | tail -6 | tee gpurun_out/${T}_memcheck.txt
