python -m pytest tests/test_gpu_deflate.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r75_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r75_pytest.log
tail -3 gpurun_out/r75_pytest.log
python bench.py --workload deflate --no-cpu > gpurun_out/r75_bench_deflate.json 2> gpurun_out/r75_bench_deflate.err; grep -o '"e2e": {[^}]*}' gpurun_out/r75_bench_deflate.json
