#!/bin/bash
# deflate: lane-per-16-bytes kernel against the 64-byte-block kernel, same library, same box
set -u
mkdir -p gpurun_out
T=${1:-r02o}
for rep in 1 2; do for v in 1 0; do
  echo -n "FDB_DEFLATE_LANE16=$v  "; FDB_DEFLATE_LANE16=$v timeout 600 python tools/gpu_lib_variants.py --one fdeflate_b200/libfdeflate_b200.so 2>&1 | tail -1
done; done | tee gpurun_out/${T}_deflate_ab.txt
echo "== pytest gpu (parity)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
