python bench.py --workload deflate --steps 3 --warmup 2 --no-cpu > gpurun_out/r57_bench_deflate.json 2> gpurun_out/r57_bench_deflate.err
echo "$(grep -o '"value": [0-9.]*, "unit": "GB/s", "n_gpus"' gpurun_out/r57_bench_deflate.json) $(grep -o '"e2e": {[^}]*}' gpurun_out/r57_bench_deflate.json)"; tail -1 gpurun_out/r57_bench_deflate.err | cut -c1-200
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r57_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r57_pytest.log
tail -3 gpurun_out/r57_pytest.log
