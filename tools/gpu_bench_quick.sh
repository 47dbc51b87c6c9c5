#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02b}
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','inflate_gbs','deflate_gbs')})
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['per_kernel'])
print('e2e', d['e2e']['value'], d['e2e'].get('frac_of_link'))
print('sweep', d['sweep']['inflate_gbs'], d['sweep']['deflate_gbs'], d['sweep']['roofline_frac_per_gpu'])
PY
