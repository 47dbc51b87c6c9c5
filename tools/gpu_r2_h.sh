#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02h}
echo "== pytest gpu (parity subset)"; timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_png.py tests/test_streaming.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
echo "== k3 2048"; timeout 400 python tools/gpu_k3_speed.py 2048 2>&1 | tail -3 | tee gpurun_out/${T}_k3.txt
echo "== k3 4096"; timeout 600 python tools/gpu_k3_speed.py 4096 2>&1 | tail -3 | tee -a gpurun_out/${T}_k3.txt
