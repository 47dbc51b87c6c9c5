#!/bin/bash
# Round-2 run e: K4 micro-variants A/B, stored-deflate kernel with bulk async copies (speed, ncu), GPU tests.
set -u
mkdir -p gpurun_out
T=${1:-r02e}
echo "== variants"; timeout 900 python tools/gpu_lib_variants.py ab/base.so ab/cur.so ab/k4adv.so ab/k4pre.so ab/k4both.so 2>&1 | tee gpurun_out/${T}_variants.txt
echo "== stored"; timeout 300 python tools/gpu_stored_speed.py 4096 2>&1 | tee gpurun_out/${T}_stored.txt
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.txt
echo "== full: deflate_stored_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deflate_stored_kernel -s 1 -c 1 -f -o gpurun_out/${T}_deflate_stored python tools/gpu_stored_speed.py 4096 > gpurun_out/${T}_deflate_stored.log 2>&1
tail -2 gpurun_out/${T}_deflate_stored.log
