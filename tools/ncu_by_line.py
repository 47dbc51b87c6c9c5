#!/usr/bin/env python3
"""Join an ncu SASS source page (per-instruction executed counts / stall samples) with nvdisasm line
info of the in-tree library, and print the hottest CUDA source lines of one kernel.

    tools/ncu_by_line.py gpurun_out/x.ncu-rep inflate_uf_kernel [top]
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "fdeflate_b200" / "libfdeflate_b200.so")], cwd=tmp,
                   capture_output=True)
    cubin = next(tmp.glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    # per-function: list of (line label) in instruction order
    func = None
    cur = ("?", 0)
    order = defaultdict(list)
    for ln in dis.splitlines():
        m = re.search(r"\.text\.(\S+):", ln) or re.search(r"//-+ \.text\.(\S+) -+", ln)
        if m:
            func = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if func and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            order[func].append(cur)
    fn = [f for f in order if kernel in f]
    if not fn:
        print("kernel not found in", list(order)[:10])
        return
    lines = order[fn[0]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ie, it, isamp = ix["Instructions Executed"], ix["Thread Instructions Executed"], ix["# Samples"]
    iwf, iwfi = ix.get("L1 Wavefronts Shared"), ix.get("L1 Wavefronts Shared Ideal")
    agg = defaultdict(lambda: [0, 0, 0, 0, 0])
    body = rows[2:]
    if len(body) != len(lines):
        print(f"warning: {len(body)} ncu instructions vs {len(lines)} disassembled")
    tot = 0
    for k, r in enumerate(body):
        if k >= len(lines) or len(r) <= ie or not r[ie].isdigit():
            continue
        a = agg[lines[k]]
        a[0] += int(r[ie])
        a[1] += int(r[it])
        a[2] += int(r[isamp]) if r[isamp].isdigit() else 0
        if iwf is not None and r[iwf].isdigit():
            a[3] += int(r[iwf])
            a[4] += int(r[iwfi]) if r[iwfi].isdigit() else 0
        tot += int(r[ie])
    tsamp = sum(a[2] for a in agg.values())
    print(f"total warp instructions {tot:,}  samples {tsamp:,}")
    srcs = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            p = list(ROOT.rglob(f))
            srcs[f] = p[0].read_text().splitlines() if p else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print(f"{a[0] / tot * 100:5.1f}% inst  {a[2] / max(tsamp, 1) * 100:5.1f}% samp  thr/inst {a[1] / max(a[0], 1):5.1f}  {f}:{l:<4} {text}")
    twf = sum(a[3] for a in agg.values())
    if twf:
        print(f"shared-memory wavefronts {twf:,} (ideal {sum(a[4] for a in agg.values()):,}), by source line:")
        for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][3])[:16]:
            text = srcs.get(f, [""] * l)[l - 1].strip()[:80] if f in srcs and 0 < l <= len(srcs[f]) else ""
            print(f"{a[3] / twf * 100:5.1f}% wavefronts  x{a[3] / max(a[4], 1):4.2f} of ideal  {f}:{l:<4} {text}")


if __name__ == "__main__":
    main()
