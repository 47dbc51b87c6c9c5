#!/bin/bash
# A/B of library builds on one box (bench tiles, device-resident), then the GPU parity tests on the in-tree library
set -u
mkdir -p gpurun_out
T=$1; shift
timeout 900 python tools/gpu_lib_variants.py "$@" "$@" 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/${T}_variants.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.txt
