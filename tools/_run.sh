python tools/sweep_inflate.py --streams 65536 --cfgs=-2,14 --steps 5 > gpurun_out/r43_sweep.log 2>&1
cat gpurun_out/r43_sweep.log
python -m pytest tests/test_gpu_inflate.py -x -q -m gpu > gpurun_out/r43_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r43_pytest.log
tail -3 gpurun_out/r43_pytest.log
