CZ_MATCH_LINKS=2 python bench.py --workload deflate --mib 1024 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r54_deflate_l2.json 2> gpurun_out/r54_deflate_l2.err
echo "links2: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r54_deflate_l2.json)"
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r54_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r54_pytest.log
tail -3 gpurun_out/r54_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r54_deflate_launches.csv python bench.py --workload deflate --mib 1024 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r54_ncu.log 2>&1
tail -1 gpurun_out/r54_ncu.log | cut -c1-100
