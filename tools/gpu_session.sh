#!/bin/bash
# One GPU session: parity tests, bench, launch list and a full ncu capture of the tensor-core kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
cat gpurun_out/bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
timeout 300 python tools/tc_time.py 1 1000000 10; timeout 300 python tools/tc_time.py 3 1000000 10
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tc -s 1 -c 1 -f -o gpurun_out/tc_prof python tools/tc_time.py 3 1000000 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
