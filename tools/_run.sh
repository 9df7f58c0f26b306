python bench.py --workload deflate --no-cpu > gpurun_out/r76_side.json 2> gpurun_out/r76_side.err; echo "side: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r76_side.json | head -2 | tr '\n' ' ')"
CZ_NO_SIDE_STREAM=1 python bench.py --workload deflate --no-cpu > gpurun_out/r76_noside.json 2> gpurun_out/r76_noside.err; echo "no side: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r76_noside.json | head -2 | tr '\n' ' ')"
python -m pytest tests/test_gpu_deflate.py -x -q -m gpu > gpurun_out/r76_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r76_pytest.log
tail -2 gpurun_out/r76_pytest.log
