#!/bin/bash
# ncu evidence on the GPU box: launch list + full captures of the two hot kernels. Output -> gpurun_out/
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
tail -3 gpurun_out/${TAG}_launches.log
echo "== full: inflate_uf_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inflate_uf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_inflate_uf $CMD > gpurun_out/${TAG}_inflate_uf.log 2>&1
tail -2 gpurun_out/${TAG}_inflate_uf.log
echo "== full: deflate_uf_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_uf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_deflate_uf $CMD > gpurun_out/${TAG}_deflate_uf.log 2>&1
tail -2 gpurun_out/${TAG}_deflate_uf.log
ls -la gpurun_out/
