#!/usr/bin/env python
"""Attribute ncu SASS-level instruction counts to source lines.

usage: sass_lines.py <ncu --page source --csv export> <nvdisasm --print-line-info output> <kernel mangled name> [top]

The ncu CLI exports per-SASS-instruction counters; nvdisasm gives the source line of every
SASS offset.  Joining them by offset gives executed warp-instructions per source line
(inlined code is attributed to the innermost line) and per inline call stack root line.
"""
import csv
import re
import sys
from collections import defaultdict

ncu_csv, sass, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
isrc = hdr.index("Source")
insts = [(int(r[ia], 16), int(r[ii]), int(r[isamp]), r[isrc]) for r in rows[2:] if len(r) > ii and r[ia].startswith("0x")]
base = insts[0][0]
# nvdisasm: track current file/line (innermost) and "inlined at" root
cur = None
root = None
line_of = {}
root_of = {}
infn = False
for ln in open(sass):
    if ln.startswith(".text."):
        infn = ln.strip() == f".text.{kernel}:"
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        rest = m.group(3)
        if "inlined at" in rest:
            m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
            root = (m2[-1][0].split("/")[-1], int(m2[-1][1])) if m2 else cur
        else:
            root = cur
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        off = int(m.group(1), 16)
        line_of[off] = cur
        root_of[off] = root
tot = sum(i[1] for i in insts)
by_line = defaultdict(lambda: [0, 0])
by_root = defaultdict(lambda: [0, 0])
by_op = defaultdict(int)
for addr, n, s, src in insts:
    off = addr - base
    by_line[line_of.get(off)][0] += n
    by_line[line_of.get(off)][1] += s
    by_root[root_of.get(off)][0] += n
    by_root[root_of.get(off)][1] += s
    by_op[src.split()[0].split(".")[0] if not src.strip().startswith("@") else src.split()[1].split(".")[0]] += n
print(f"total warp instructions {tot}")
print("-- by innermost source line")
for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(k):40s} {v[0]:14d} {100*v[0]/tot:6.2f}%  samples {v[1]}")
print("-- by inline root line (line of the kernel body)")
for k, v in sorted(by_root.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(k):40s} {v[0]:14d} {100*v[0]/tot:6.2f}%  samples {v[1]}")
print("-- by opcode")
for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{k:12s} {v:14d} {100*v/tot:6.2f}%")
