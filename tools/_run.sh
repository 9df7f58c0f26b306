python -m pytest tests/test_gpu_inflate.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r90_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r90_pytest.log
tail -3 gpurun_out/r90_pytest.log
timeout 600 python tools/fuzz_gpu.py 1000 10 > gpurun_out/r90_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/r90_fuzz.log; tail -4 gpurun_out/r90_fuzz.log
