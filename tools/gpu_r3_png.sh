#!/bin/bash
# the filter fused into the ultra-fast encoder: GPU tests of the PNG paths, then the device-resident speed of both ways
set -u
mkdir -p gpurun_out
T=${1:-r03p}
echo "== pytest gpu (png, host api)"; timeout 900 python -m pytest tests/test_png.py tests/test_host_api.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
echo "== speed"; timeout 600 python tools/gpu_png_fused_speed.py 4096 2>&1 | tail -6 | tee gpurun_out/${T}_png_fused_speed.txt
