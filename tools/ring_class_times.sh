B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-svd --no-generate --no-graph"
for cfg in "DSB_RING_SMEM_KB=64" "DSB_RING_SMEM_KB=113" "DSB_RING_SMEM_KB=140" "DSB_RING_NPP=1 DSB_RING_SMEM_KB=96"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg ncu --metrics gpu__time_duration.sum,launch__shared_mem_per_block_dynamic --clock-control none -k regex:ringfft --csv --log-file gpurun_out/ringcls_$tag.csv $B > /dev/null 2>&1
done
