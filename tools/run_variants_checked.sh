#!/bin/bash
# Time every variant library (tools/build_variants.py) and run the tensor-core parity tests against it (dev builds: d = 4 only).
for lib in nmma_b200/lib/variants/lib_*.so; do
  echo "== $lib"
  NMMA_B200_LIB=$PWD/$lib timeout 120 python tools/tc_time.py 3 1000000 10 2>&1 | grep -v KNtheta | tail -1
  NMMA_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bu2019lm or fp32_accuracy or dynamic_range or averaged or n_coeff" 2>&1 | tail -2
done
