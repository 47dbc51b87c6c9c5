#!/bin/bash
# A/B of two builds of the library on the same box: ab/base.so vs ab/new.so, alternating, device-resident numbers only
set -u
for round in 1 2 3; do
  for v in base new; do
    cp ab/$v.so fdeflate_b200/libfdeflate_b200.so
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['inflate_gbs'], d['deflate_gbs'])"
  done
done
