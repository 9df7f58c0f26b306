python -m pytest tests -x -q -m gpu > gpurun_out/r85_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r85_pytest.log
tail -3 gpurun_out/r85_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r85_smoke.log 2>&1; tail -2 gpurun_out/r85_smoke.log
python bench.py > gpurun_out/r85_bench.json 2> gpurun_out/r85_bench.err; cut -c1-200 gpurun_out/r85_bench.json
python bench.py --workload deflate > gpurun_out/r85_bench_deflate.json 2> gpurun_out/r85_bench_deflate.err; cut -c1-200 gpurun_out/r85_bench_deflate.json
python bench.py --impl reference > gpurun_out/r85_ref.json 2> gpurun_out/r85_ref.err; cut -c1-160 gpurun_out/r85_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r85_deflate_launches.csv python bench.py --workload deflate --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r85_ncu_deflate.log 2>&1
