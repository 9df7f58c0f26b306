python -m pytest tests/test_gpu_deflate.py -x -q -m gpu -k "cfg5" --durations=3 > gpurun_out/r67_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r67_pytest.log
tail -12 gpurun_out/r67_pytest.log
