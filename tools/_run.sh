python bench.py --workload deflate --no-cpu > gpurun_out/r83_split.json 2> gpurun_out/r83_split.err; echo "split: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r83_split.json | head -2 | tr '\n' ' ')"; tail -1 gpurun_out/r83_split.err | cut -c1-200
CZ_NO_CHAIN_SPLIT=1 python bench.py --workload deflate --no-cpu --no-e2e > gpurun_out/r83_nosplit.json 2> gpurun_out/r83_nosplit.err; echo "no split: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r83_nosplit.json | head -1)"
python -m pytest tests/test_gpu_deflate.py -x -q -m gpu > gpurun_out/r83_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r83_pytest.log
tail -2 gpurun_out/r83_pytest.log
