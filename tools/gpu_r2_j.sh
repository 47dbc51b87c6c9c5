#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02j}
echo "== pytest gpu (parity)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
echo "== sweep 8192"; timeout 900 python tools/gpu_sweep.py 8192 2>&1 | tail -3 | tee gpurun_out/${T}_sweep.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 3000 gpurun_out/${T}_bench.json
