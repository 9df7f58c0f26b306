python bench.py > gpurun_out/r62_bench.json 2> gpurun_out/r62_bench.err; cut -c1-220 gpurun_out/r62_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r62_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r62_ncu.log 2>&1
python -m pytest tests/test_gpu_inflate.py -x -q -m gpu > gpurun_out/r62_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r62_pytest.log
tail -2 gpurun_out/r62_pytest.log
