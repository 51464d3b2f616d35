#!/bin/bash
# One GPU session: bench (both arms), accuracy print, launch list and a full ncu capture of the tensor-core kernel.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err; echo "bench rc=$?"
cat gpurun_out/r02b_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_ref.json 2>/dev/null; cat gpurun_out/r02b_bench_ref.json
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fp32_accuracy" 2>&1 | grep "err" > gpurun_out/r02b_accuracy.txt; cat gpurun_out/r02b_accuracy.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tc -s 1 -c 1 -f -o gpurun_out/r02b_tc_prof python tools/tc_time.py 3 1000000 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -8
