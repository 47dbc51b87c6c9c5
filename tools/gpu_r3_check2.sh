#!/bin/bash
# after the source-object refactor of the encoder: A/B against the build before it, the whole GPU suite, memcheck over
# the fused PNG encode test
set -u
mkdir -p gpurun_out
T=${1:-r03q}
bash tools/gpu_ab_list.sh ${T} ab/t3.so ab/u1.so
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
echo "== memcheck: fused png encode, pinned direct output"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_png.py tests/test_host_api.py -m gpu -x -q -k "fused or pinned" > gpurun_out/${T}_sanitizer_memcheck_png_fused.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_memcheck_png_fused.log | head -5
