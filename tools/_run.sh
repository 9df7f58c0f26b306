python bench.py --workload deflate --no-cpu > gpurun_out/r82_bench_deflate.json 2> gpurun_out/r82_bench_deflate.err; cut -c1-200 gpurun_out/r82_bench_deflate.json; grep -o '"e2e": {[^}]*}' gpurun_out/r82_bench_deflate.json
python -m pytest tests/test_gpu_deflate.py -x -q -m gpu > gpurun_out/r82_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r82_pytest.log
tail -2 gpurun_out/r82_pytest.log
