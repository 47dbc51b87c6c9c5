#!/bin/bash
# fourth session, evidence with the final build: the whole GPU suite, smoke, randomised parity (one seed), bench line,
# reference arm, launch list of the bench command, ncu --set full of the encoder, memcheck over smoke and the deflate tests
set -u
mkdir -p gpurun_out
T=${1:-r04z}
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== fuzz"; timeout 900 python tools/gpu_fuzz_uf.py 3000 4 2>&1 | tail -2 | tee gpurun_out/${T}_fuzz.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; tail -c 200 gpurun_out/${T}_bench_reference_arm.json
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --sweep-streams 256"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_launches.log 2>&1; tail -1 gpurun_out/${T}_launches.log | cut -c1-200
echo "== ncu full: deflate_ufb_kernel, inflate_uf_kernel"
bash tools/gpu_prof_one.sh ${T}_deflate_ufb_kernel deflate_ufb_kernel
bash tools/gpu_prof_one.sh ${T}_inflate_uf_kernel inflate_uf_kernel
echo "== memcheck smoke + deflate tests"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_memcheck_smoke.log 2>&1
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/${T}_sanitizer_memcheck_smoke.log | head -3
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deflate or segment_by_segment or fast_path or span_by_span" > gpurun_out/${T}_sanitizer_memcheck_deflate_tests.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_memcheck_deflate_tests.log | head -3
