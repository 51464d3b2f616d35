timeout 300 python tools/tc_time.py 3 1000000 10
TC_TIME_CONFIG=c3 timeout 300 python tools/tc_time.py 3 1000000 10
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
NMMA_B200_LIB=$PWD/nmma_b200/lib/variants/lib_n32b4s2k16.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_vectors.py -m gpu -x -q > gpurun_out/pytest_gpu_variant.log 2>&1; echo "variant pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_variant.log
