#!/bin/bash
# --set full of the dominant ring launch (equatorial rings of the nside-256 bucket, L = 1024)
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-svd"
ncu --set full --clock-control none --import-source on -k regex:ringfft_kernel --launch-skip 64 -c 1 -f -o /tmp/ringmain $B > $OUT/ncu_ringmain.log 2>&1
python profiles/ncu_summary.py /tmp/ringmain.ncu-rep > $OUT/r01_ncu_ring_main.txt 2>&1
ncu -i /tmp/ringmain.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for k in ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_fp64.sum','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','launch__grid_size','launch__waves_per_multiprocessor','sm__maximum_warps_per_active_cycle_pct'):
        if k in h: print(k, r[h.index(k)])
" >> $OUT/r01_ncu_ring_main.txt
# source-level hot spots: top 25 SASS lines by stall samples, mapped to source
ncu -i /tmp/ringmain.ncu-rep --page source --csv 2>/dev/null > /tmp/ring_source.csv
python - <<'PY' >> $OUT/r01_ncu_ring_main.txt 2>&1
import csv
rows = list(csv.reader(open('/tmp/ring_source.csv')))
hdr = next(r for r in rows if 'Address' in r)
i0 = rows.index(hdr)
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ins = [(int(r[isamp] or 0), int(r[iex] or 0), r[ia], r[isrc]) for r in rows[i0 + 1:] if len(r) > isamp and r[ia].startswith('0x')]
tot = sum(i[0] for i in ins) or 1
print('total samples', tot, 'instructions', len(ins))
# opcode histogram by samples and by executed count
from collections import Counter
bys, bye = Counter(), Counter()
for smp, ex, a, src in ins:
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    bys[op] += smp; bye[op] += ex
print('samples by opcode:', [(k, round(100 * v / tot, 1)) for k, v in bys.most_common(14)])
te = sum(bye.values()) or 1
print('executed by opcode:', [(k, round(100 * v / te, 1)) for k, v in bye.most_common(14)])
for smp, ex, a, src in sorted(ins, reverse=True)[:25]:
    print(f'{100 * smp / tot:5.1f}%  {ex:>10d}  {a}  {src[:90]}')
PY
