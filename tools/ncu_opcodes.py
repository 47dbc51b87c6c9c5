#!/usr/bin/env python3
"""Dynamic SASS opcode mix of a kernel from an ncu report's source page: executed warp instructions per opcode.
    tools/ncu_opcodes.py gpurun_out/x.ncu-rep [top]"""
import csv, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
isrc, ie = ix["Source"], ix["Instructions Executed"]
c = Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit(): continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", r[isrc])
    if not m: continue
    op = m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("IMAD", "SHF", "LOP3", "IADD3")) and "." in m.group(1) else "")
    c[op] += int(r[ie]); tot += int(r[ie])
print(f"total {tot:,}")
for op, n in c.most_common(top): print(f"{n/tot*100:5.1f}%  {op}")
