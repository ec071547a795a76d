#!/bin/bash
# Round-2 profiles of one bench step (sht_iter = 2): launch list with DRAM traffic, and --set full
# summaries of the ring kernel's largest launch, the a(0) analysis, the two refinement contractions,
# the alias fold and the update kernel of the nside-256 bucket.  Leaves text/csv in gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-svd --no-generate --no-graph"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/r02_launches_bench_steps1.csv $B > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"ringfft_kernel<float, .int.10, .int.0>" --launch-skip 2 -c 1 -f -o /tmp/ringmain $B > $OUT/ncu_ringmain.log 2>&1
python profiles/ncu_summary.py /tmp/ringmain.ncu-rep > $OUT/r02_ncu_ring_main.txt 2>&1
python profiles/ncu_hot.py /tmp/ringmain.ncu-rep >> $OUT/r02_ncu_ring_main.txt 2>&1
# legendre launches per step: 3 buckets x (a0, [synthesis-direction, cap analysis] x 2); the nside-256
# bucket is the second one: launches 20 (a0), 21 (synthesis direction), 22 (cap analysis) counting the warm-up
ncu --set full --clock-control none --import-source on -k regex:legendre_tc_kernel --launch-skip 20 -c 3 -f -o /tmp/leg $B > $OUT/ncu_leg.log 2>&1
python profiles/ncu_summary.py /tmp/leg.ncu-rep > $OUT/r02_ncu_legendre.txt 2>&1
python profiles/ncu_hot.py /tmp/leg.ncu-rep >> $OUT/r02_ncu_legendre.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"alias_fold_bins|refine_update" --launch-skip 15 -c 3 -f -o /tmp/ref $B > $OUT/ncu_ref.log 2>&1
python profiles/ncu_summary.py /tmp/ref.ncu-rep > $OUT/r02_ncu_refine.txt 2>&1
ls -la $OUT | tail -8
