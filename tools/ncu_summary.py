#!/usr/bin/env python3
"""Text summary of one ncu report (`--set full --import-source on`), written under profiles/:
headline metrics, the hottest CUDA source lines (executed warp instructions, sampled stalls, shared-memory wavefronts
and their excess over the ideal) and the dynamic opcode mix.
    tools/ncu_summary.py gpurun_out/x.ncu-rep [min_percent] > profiles/x_ncu_summary.txt"""
import csv
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, v = rows[0], rows[1], rows[2]
print(v[h.index("Kernel Name")] if "Kernel Name" in h else rep)
for m in METRICS:
    if m in h:
        print(f"{m:92s} {v[h.index(m)]:>16s} {units[h.index(m)]}")

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, res, ops = None, None, [], Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) < len(hdr) - 2:
        continue
    ie = r[hdr.index("Instructions Executed")]
    if not ie.isdigit():
        continue
    if r[2] == "-":  # a CUDA source line
        res.append((cur, int(r[0]), int(ie), int(r[hdr.index("Thread Instructions Executed")] or 0),
                    int(r[hdr.index("L1 Wavefronts Shared")] or 0), int(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0),
                    int(r[hdr.index("# Samples")] or 0), r[1][:84]))
    else:  # a SASS instruction under it
        m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", r[3])
        if m:
            ops[m.group(1)] += int(ie)
T = sum(x[2] for x in res) or 1
W = sum(x[4] for x in res) or 1
S = sum(x[6] for x in res) or 1
print(f"\nhottest source lines (>= {thr} % of the executed warp instructions, of the stall samples, or >= 1 % of the shared-memory wavefronts)")
print("file:line                 inst%  smp%   thr/inst  smem-wf%  x ideal | source")
for f, l, ie, tie, wf, wfi, smp, src in res:
    if ie / T * 100 >= thr or wf / W * 100 >= 1.0 or smp / S * 100 >= thr:
        print(f"{f}:{l:<5d} {ie / T * 100:6.2f} {smp / S * 100:6.2f} {tie / max(ie, 1):9.1f} {wf / W * 100:8.1f} {wf / max(wfi, 1):8.2f} | {src}")
O = sum(ops.values()) or 1
print("\ndynamic opcode mix (executed warp instructions)")
print("  ".join(f"{op} {n / O * 100:.1f}%" for op, n in ops.most_common(18)))
