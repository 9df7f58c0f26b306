for l in 2 1; do
CZ_MATCH_LINKS=$l python bench.py --workload deflate --mib 1024 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r51_deflate_l$l.json 2> gpurun_out/r51_deflate_l$l.err
cut -c1-200 gpurun_out/r51_deflate_l$l.json; tail -2 gpurun_out/r51_deflate_l$l.err
done
python -m pytest tests/test_gpu_deflate.py -x -q -m gpu > gpurun_out/r51_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r51_pytest.log
tail -3 gpurun_out/r51_pytest.log
