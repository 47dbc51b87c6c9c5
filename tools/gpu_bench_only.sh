#!/bin/bash
# quick GPU check: GPU parity tests for the kernels + short bench (no CPU baseline)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'inflate',d['inflate_gbs'],'deflate',d['deflate_gbs'],'e2e',d['e2e']['value'],d['roofline']['per_kernel'])
PY
