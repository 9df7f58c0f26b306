python -m pytest tests -x -q -m gpu > gpurun_out/r97_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r97_pytest.log
tail -3 gpurun_out/r97_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r97_smoke.log 2>&1; tail -1 gpurun_out/r97_smoke.log
