#!/bin/bash
# Runs on the GPU box via gpurun: smoke, GPU parity tests, a short bench. Logs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err ; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
