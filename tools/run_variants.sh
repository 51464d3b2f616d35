#!/bin/bash
# Time every variant library built by tools/build_variants.py: tools/run_variants.sh [PATH] [CONFIG] [POINTS]
for lib in nmma_b200/lib/variants/lib_*.so; do
  echo "== $lib"
  NMMA_B200_LIB=$PWD/$lib TC_TIME_CONFIG=${2:-c2} timeout 90 python tools/tc_time.py ${1:-3} ${3:-1000000} 10 2>&1 | grep -v KNtheta | tail -1
done
