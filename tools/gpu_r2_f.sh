#!/bin/bash
# Round-2 run f: general-inflate (K3) table-size variants, then the bench line of the current build.
set -u
mkdir -p gpurun_out
T=${1:-r02f}
for v in k3_12 k3_11 k3_10 k3_10m; do timeout 400 python tools/gpu_k3_speed.py 2048 ab/$v.so 2>&1 | tail -3; done | tee gpurun_out/${T}_k3_variants.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['inflate_gbs'], d['deflate_gbs'], d['roofline']['per_kernel'], d['e2e']['value'], d['clocks'], d['sweep']['inflate_gbs'], d['sweep']['deflate_gbs'])"
