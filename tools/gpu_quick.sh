#!/bin/bash
# Quick GPU check: device timing of the tensor-core (3) throughput kernel first, then the parity tests.
mkdir -p gpurun_out
timeout 300 python tools/tc_time.py 3 1000000 10
TC_TIME_CONFIG=c3 timeout 300 python tools/tc_time.py 3 1000000 10
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
