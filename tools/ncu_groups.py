#!/usr/bin/env python3
"""Instruction / stall-sample shares of a kernel grouped by source regions delimited by marker strings.
   tools/ncu_groups.py <rep> <kernel> <source.cuh> <units_per_launch> marker1 marker2 ...   (markers in file order)"""
import re, subprocess, sys
from pathlib import Path
rep, kernel, src, units = sys.argv[1], sys.argv[2], Path(sys.argv[3]), float(sys.argv[4])
marks = sys.argv[5:]
lines = src.read_text().splitlines()
def find(s):
    return [i + 1 for i, l in enumerate(lines) if s in l][0]
bounds = [(m, find(m)) for m in marks]
out = subprocess.run([sys.executable, str(Path(__file__).parent / "ncu_by_line.py"), rep, kernel, "100000"], capture_output=True, text=True).stdout
tot = {m: [0.0, 0.0] for m, _ in bounds}
tot["(before first marker)"] = [0.0, 0.0]
other = {}
T = 0
for ln in out.splitlines():
    m0 = re.match(r"total warp instructions ([\d,]+)", ln)
    if m0:
        T = int(m0.group(1).replace(",", ""))
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% samp.*?thr/inst\s+[\d.]+\s+(\S+):(\d+)", ln)
    if not m:
        continue
    pct, sp, f, l = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
    if f == src.name:
        name = "(before first marker)"
        for mk, a in bounds:
            if l >= a:
                name = mk
        tot[name][0] += pct
        tot[name][1] += sp
    else:
        o = other.setdefault(f, [0.0, 0.0])
        o[0] += pct
        o[1] += sp
per = T / units / 100.0
print(f"total warp instructions {T:,} = {T / units:.0f} per unit")
for k, v in tot.items():
    print(f"{v[0]:5.1f}% inst ({v[0] * per:6.0f}/unit) {v[1]:5.1f}% samp  {k[:70]}")
for k, v in sorted(other.items(), key=lambda x: -x[1][0]):
    print(f"{v[0]:5.1f}% inst ({v[0] * per:6.0f}/unit) {v[1]:5.1f}% samp  {k}")
