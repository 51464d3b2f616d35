#!/bin/bash
# Quick GPU check: parity tests, then device timing of the FFMA (1) and tensor-core (3) throughput kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/tc_time.py 1 1000000 10; timeout 300 python tools/tc_time.py 3 1000000 10
[ -x tools/fp64_rate ] && timeout 120 tools/fp64_rate > gpurun_out/fp64_rate.txt 2>&1; cat gpurun_out/fp64_rate.txt
