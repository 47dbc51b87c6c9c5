#!/bin/bash
# ultra-fast deflate writing pinned host output itself (FDB_DIRECT_OUT) against the payload-copy path: the test of the
# direct path, then the bench's e2e number both ways, alternating
set -u
mkdir -p gpurun_out
T=${1:-r03d}
echo "== test"; timeout 600 python -m pytest tests/test_host_api.py tests/test_slot_layouts.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
for rep in 1 2; do for v in 0 1; do
  echo -n "FDB_DIRECT_OUT=$v  "
  FDB_DIRECT_OUT=$v timeout 600 python bench.py --sweep-streams 0 --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('e2e %.3f GB/s  (link %.1f GB/s per direction, %.3f of it)  value %.1f' % (e['value'], e['link']['gbs_per_direction_all_gpus'], e['frac_of_link'], d['value']))"
done; done | tee gpurun_out/${T}_direct_out.txt
