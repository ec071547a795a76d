#!/bin/bash
# Profiles one bench step on the GPU box and leaves only small text/csv summaries in gpurun_out/
# (the .ncu-rep files are exported on the box and deleted: gpurun_out is capped at 64 MiB).
set -u
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-svd"
# (1) launch list with durations and DRAM traffic of every kernel of warm-up + 1 step
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv \
    --log-file $OUT/r01_launches.csv $B > $OUT/ncu_launches.log 2>&1
# (2) --set full of the nside-256 bucket's kernels (second step: skip the warm-up's launches)
ncu --set full --clock-control none --import-source on -k regex:ringfft_kernel --launch-skip 65 -c 13 -f -o /tmp/ring $B > $OUT/ncu_ring.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:legendre_tc --launch-skip 5 -c 1 -f -o /tmp/leg $B > $OUT/ncu_leg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_mmajor --launch-skip 5 -c 1 -f -o /tmp/pack $B > $OUT/ncu_pack.log 2>&1
for k in ring leg pack; do
  python profiles/ncu_summary.py /tmp/$k.ncu-rep > $OUT/r01_ncu_$k.txt 2>&1
done
ls -la $OUT
