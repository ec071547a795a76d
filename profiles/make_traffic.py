#!/usr/bin/env python
"""Per-kernel time shares and DRAM bytes of the timed step from an ncu launch list.

usage: make_traffic.py profiles/r02_launches_bench_steps1.csv profiles/r02_traffic.json [buckets=3]

The list comes from
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-svd --no-generate --no-graph
(tools/profile_r02.sh).  A step ends with one pack kernel per nside bucket: the timed step is everything
after the pack kernel that closes the warm-up step.
"""
import collections
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(int(r[ii]), {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
launches = [per[k] for k in sorted(per)]
packs = [i for i, l in enumerate(launches) if "pack_mmajor" in l["name"]]
start = packs[-nb - 1] + 1
step = launches[start:]
agg = collections.OrderedDict()
prev = ""
for l in step:
    n = l["name"].split("(")[0].replace("void ", "").replace("dsb::", "").split("<")[0]
    # the contraction right after a bucket's ring launches is the a(0) analysis (stage 2 of the bench line);
    # the others belong to the Jacobi refinement
    if n == "legendre_tc_kernel" and prev != "ringfft_kernel":
        n = "legendre_tc_kernel (refinement)"
    prev = n.split(" ")[0] if n.startswith("legendre") and prev == "ringfft_kernel" else n
    a = agg.setdefault(n, {"launches_per_step": 0, "ncu_time_ns_per_step": 0.0, "dram_bytes_per_step": 0.0})
    a["launches_per_step"] += 1
    a["ncu_time_ns_per_step"] += l["gpu__time_duration.sum"]
    a["dram_bytes_per_step"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
tot = sum(a["ncu_time_ns_per_step"] for a in agg.values())
for a in agg.values():
    a["share_of_step"] = a["ncu_time_ns_per_step"] / tot
agg["_source"] = (f"{src}, launches {start}-{len(launches) - 1} (= the timed step); "
                  "dram__bytes_read.sum + dram__bytes_write.sum; see profiles/make_traffic.py")
json.dump(agg, open(dst, "w"), indent=1)
print(json.dumps(agg, indent=1))
