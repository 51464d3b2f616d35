#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
cat gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
