#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02g}
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.txt
echo "== k3"; timeout 400 python tools/gpu_k3_speed.py 2048 2>&1 | tail -3 | tee gpurun_out/${T}_k3.txt
echo "== stored"; timeout 300 python tools/gpu_stored_speed.py 4096 2>&1 | tee gpurun_out/${T}_stored.txt
echo "== full: inflate_general_kernel (zlib-6 tiles)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inflate_general_kernel -s 2 -c 1 -f -o gpurun_out/${T}_k3 python tools/gpu_k3_speed.py 1024 > gpurun_out/${T}_k3_ncu.log 2>&1; tail -2 gpurun_out/${T}_k3_ncu.log
