#!/bin/bash
# Round-2 K4 experiment on the GPU box: variants A/B, GPU parity tests, ncu capture of the new inflate kernel.
set -u
mkdir -p gpurun_out
T=${1:-r02a}
echo "== variants"; timeout 900 python tools/gpu_lib_variants.py ab/base.so ab/w20.so ab/w16.so ab/w22.so ab/w24win3k.so ab/s12w15.so ab/s12w16win4k.so ab/s16w11.so 2>&1 | tee gpurun_out/${T}_variants.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.txt
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
echo "== full: inflate_uf_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^inflate_uf_kernel -s 1 -c 1 -f -o gpurun_out/${T}_inflate_uf_kernel $CMD > gpurun_out/${T}_inflate_uf_kernel.log 2>&1
tail -2 gpurun_out/${T}_inflate_uf_kernel.log
