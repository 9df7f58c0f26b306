python tools/sweep_inflate.py --streams 65536 --cfgs=-2,14 --steps 5 > gpurun_out/r93_sweep.log 2>&1; tail -1 gpurun_out/r93_sweep.log | cut -c1-120
python -m pytest tests/test_gpu_inflate.py -x -q -m gpu > gpurun_out/r93_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r93_pytest.log; tail -2 gpurun_out/r93_pytest.log
timeout 200 python tools/fuzz_gpu.py 7000 6 > gpurun_out/r93_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/r93_fuzz.log; tail -2 gpurun_out/r93_fuzz.log
