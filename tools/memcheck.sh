#!/bin/bash
# compute-sanitizer memcheck over the kernels of round 2: tools/memcheck.sh > gpurun_out/r02_memcheck.txt
run() { echo "== $*"; timeout 600 compute-sanitizer --tool memcheck "$@" 2>&1 | grep -E "ERROR SUMMARY|Invalid|path [0-9]|^C[34]" | head -8; }
run python tools/tc_time.py 3 70000 1                      # fused_tc_logl_kernel<10,1,0,0> (un-split)
run python tools/tc_time.py 3 4096 1                       # filter-split + combine_parts_kernel
run python tools/tc_time.py 5 300 1                        # latency path: coefficient mode with hidden split + backend_logl_parts_fast_kernel
TC_TIME_CONFIG=c4 run python tools/tc_time.py 4 5000 1     # fused_gp_logl_kernel (ragged last tile, tickets)
TC_TIME_CONFIG=c4 run python tools/tc_time.py 2 500 1      # coeff_gp_kernel + backend_logl_kernel
TC_TIME_CONFIG=c3 run python tools/tc_time.py 3 3000 1     # sampled-systematics / detection-limit classes (fast_log_ndtr)
echo "== averaged filters / n_coeff 7 through the coefficient mode (pytest)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -k "averaged or n_coeff" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | head -4
