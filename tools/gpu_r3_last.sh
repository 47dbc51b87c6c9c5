#!/bin/bash
# last evidence run of the session with the final build: randomised parity (two seeds), the bench line, the reference
# arm, the launch list of the bench command
set -u
mkdir -p gpurun_out
T=${1:-r03z}
echo "== fuzz"; for seed in 1 2; do timeout 900 python tools/gpu_fuzz_uf.py 3000 $seed 2>&1 | tail -2; done | tee gpurun_out/${T}_fuzz.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 700 gpurun_out/${T}_bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; tail -c 400 gpurun_out/${T}_bench_reference_arm.json
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --sweep-streams 256"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_launches.log 2>&1; tail -1 gpurun_out/${T}_launches.log | cut -c1-200
