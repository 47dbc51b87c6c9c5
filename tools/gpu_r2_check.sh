#!/bin/bash
# Round-2 check on the GPU box: GPU parity tests, smoke, the bench line (tiles + sweep), the reference arm.
set -u
mkdir -p gpurun_out
T=${1:-r02b}
nproc > gpurun_out/${T}_nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.txt
echo "== bench (ours)"; timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 6000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; tail -c 1500 gpurun_out/${T}_bench_reference.json
