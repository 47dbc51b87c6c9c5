#!/bin/bash
# evidence run with the final build of the round: GPU test suite, bench line, reference arm, launch list, ncu captures
set -u
mkdir -p gpurun_out
T=${1:-r02f}
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 1500 gpurun_out/${T}_bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; tail -c 600 gpurun_out/${T}_bench_reference_arm.json
bash tools/gpu_r2_profile.sh ${T}
