#!/bin/bash
# Round-1 final evidence run on the GPU box: bench lines (ours, reference arm), launch list, ncu --set full of the
# hot kernels, of the stored-deflate kernel and of the span / segment passes.  Output -> gpurun_out/
set -u
mkdir -p gpurun_out
T=r01v11
echo "== bench (ours)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench.json
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench_reference.json
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_launches.log 2>&1
for k in inflate_uf_kernel deflate_uf_kernel; do
  echo "== full: $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o gpurun_out/${T}_$k $CMD > gpurun_out/${T}_$k.log 2>&1
  tail -1 gpurun_out/${T}_$k.log
done
echo "== full: deflate_stored_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_stored_kernel -s 1 -c 1 -f -o gpurun_out/${T}_deflate_stored python tools/gpu_stored_speed.py 4096 > gpurun_out/${T}_deflate_stored.log 2>&1
tail -1 gpurun_out/${T}_deflate_stored.log
for k in inflate_uf_split_count_kernel inflate_uf_split_write_kernel deflate_uf_split_count_kernel deflate_uf_split_write_kernel; do
  echo "== full: $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${T}_$k python tools/gpu_sweep.py 512 16 > gpurun_out/${T}_$k.log 2>&1
  tail -1 gpurun_out/${T}_$k.log
done
python tools/gpu_sweep.py 2048 16 > gpurun_out/${T}_sweep.txt 2>&1; python tools/gpu_sweep.py 256 16 >> gpurun_out/${T}_sweep.txt 2>&1; cat gpurun_out/${T}_sweep.txt
python tools/gpu_stored_speed.py 4096 > gpurun_out/${T}_stored.txt 2>&1; cat gpurun_out/${T}_stored.txt
python tools/gpu_k3_speed.py 2048 > gpurun_out/${T}_k3.txt 2>&1; cat gpurun_out/${T}_k3.txt
ls -la gpurun_out/ | head -40
