#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02i}
echo "== pytest gpu (span tests)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "span or sweep or config5" 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
echo "== sweep records 4096"; timeout 900 python tools/gpu_sweep_records.py 4096 2>&1 | tail -12 | tee gpurun_out/${T}_sweep_records.txt
