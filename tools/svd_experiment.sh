#!/bin/bash
# diagnostic: sweep counts of the block Jacobi under different settings
for v in "" "DSB_SVD_NOSORT=1" "DSB_SVD_INNER=3"; do
  echo "=== $v"
  env $v DSB_SVD_DEBUG=1 timeout 120 python tools/bench_svd.py --only 1 2>&1 | awk '/jacobi/{n++; last=$0} /batch|Error/{print} END{print n, last}'
done
echo "=== full"
timeout 400 python tools/bench_svd.py --big 2>&1 | tail -8
