#!/bin/bash
# ncu --set full of the general inflate kernel on zlib-6 tiles: tools/gpu_prof_k3.sh <tag>
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inflate_general_kernel -s 3 -c 1 -f -o gpurun_out/$1 python tools/gpu_k3_speed.py 1024 > gpurun_out/$1.log 2>&1
tail -3 gpurun_out/$1.log
