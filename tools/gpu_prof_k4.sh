#!/bin/bash
# one --set full capture of inflate_uf_kernel (bench tiles) with the current build
set -u
mkdir -p gpurun_out
T=${1:-r02k4}
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --sweep-streams 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^${K:-inflate_uf_kernel} -s 1 -c 1 -f -o gpurun_out/${T}_${K:-inflate_uf_kernel} $CMD > gpurun_out/${T}_${K:-inflate_uf_kernel}.log 2>&1
tail -1 gpurun_out/${T}_${K:-inflate_uf_kernel}.log
