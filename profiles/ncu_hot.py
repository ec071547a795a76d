#!/usr/bin/env python
"""Hot spots of an .ncu-rep captured with --import-source on: extra pipe metrics, opcode histograms and
the 25 SASS lines with the most stall samples.  usage: ncu_hot.py file.ncu-rep"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
for r in rows[2:3]:
    for k in ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "launch__grid_size",
              "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct"):
        if k in h:
            print(k, r[h.index(k)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if "Address" in r)
i0 = rows.index(hdr)
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = [(int(r[isamp] or 0), int(r[iex] or 0), r[ia], r[isrc]) for r in rows[i0 + 1:] if len(r) > isamp and r[ia].startswith("0x")]
tot = sum(i[0] for i in ins) or 1
print("total samples", tot, "instructions", len(ins))
bys, bye = Counter(), Counter()
for smp, ex, a, s in ins:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    bys[op] += smp
    bye[op] += ex
print("samples by opcode:", [(k, round(100 * v / tot, 1)) for k, v in bys.most_common(14)])
te = sum(bye.values()) or 1
print("executed by opcode:", [(k, round(100 * v / te, 1)) for k, v in bye.most_common(14)])
for smp, ex, a, s in sorted(ins, reverse=True)[:25]:
    print(f"{100 * smp / tot:5.1f}%  {ex:>10d}  {a}  {s[:90]}")
