#!/bin/bash
# ncu --set full for one kernel: tools/gpu_prof_one.sh <tag> <kernel-regex>
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/$1 $CMD > gpurun_out/$1.log 2>&1
tail -2 gpurun_out/$1.log
