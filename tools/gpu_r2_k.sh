#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02k}
echo "== variants"; timeout 900 python tools/gpu_lib_variants.py ab/base.so ab/v1.so ab/v2.so ab/base.so ab/v1.so ab/v2.so 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/${T}_variants.txt
echo "== pytest gpu (parity)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
