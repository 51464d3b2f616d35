#!/bin/bash
# Time every variant library built by tools/build_variants.py on the bench workload (path 3).
for lib in nmma_b200/lib/variants/lib_*.so; do
  echo "== $lib"
  NMMA_B200_LIB=$PWD/$lib timeout 90 python tools/tc_time.py 3 1000000 10 2>&1 | grep -v KNtheta | tail -1
done
