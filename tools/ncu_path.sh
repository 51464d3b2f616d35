#!/bin/bash
# ncu --set full capture of one kernel path on the bench workload: tools/ncu_path.sh PATH REGEX OUTNAME [POINTS]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/$3 python tools/tc_time.py $1 ${4:-1000000} 1 > gpurun_out/$3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/$3.log
