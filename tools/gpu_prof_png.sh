#!/bin/bash
# ncu --set full of the PNG row-filter kernels: tools/gpu_prof_png.sh <tag>
set -u
mkdir -p gpurun_out
for k in png_unfilter_kernel png_filter_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/$1_$k python tools/gpu_png_speed.py 4096 > gpurun_out/$1_$k.log 2>&1
  tail -2 gpurun_out/$1_$k.log
done
