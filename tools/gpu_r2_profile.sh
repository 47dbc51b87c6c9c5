#!/bin/bash
# Round-2 evidence run: launch list of the bench command, --set full captures of the two hot kernels and of the
# general kernel on zlib-6 tiles (2048 streams), with the final build.
set -u
mkdir -p gpurun_out
T=${1:-r02v2}
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --sweep-streams 256"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_launches.log 2>&1; tail -1 gpurun_out/${T}_launches.log | cut -c1-300
for k in inflate_uf_kernel deflate_ufb_kernel; do
  echo "== full: $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o gpurun_out/${T}_$k $CMD > gpurun_out/${T}_$k.log 2>&1
  tail -1 gpurun_out/${T}_$k.log
done
echo "== full: inflate_general_kernel, 2048 zlib-6 tiles"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inflate_general_kernel -s 2 -c 1 -f -o gpurun_out/${T}_k3 python tools/gpu_k3_speed.py 2048 > gpurun_out/${T}_k3_ncu.log 2>&1; tail -1 gpurun_out/${T}_k3_ncu.log
