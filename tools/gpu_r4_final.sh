#!/bin/bash
# fourth session: the whole GPU suite, smoke, the bench line and the reference arm with the build of the committed source
set -u
mkdir -p gpurun_out
T=${1:-r04}
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; tail -c 300 gpurun_out/${T}_bench_reference_arm.json
