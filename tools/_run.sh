set -x
python -m pytest tests/test_gpu_inflate.py -x -q -m gpu -k "cta or fuzz" > gpurun_out/r41_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r41_pytest.log
tail -5 gpurun_out/r41_pytest.log
python tools/sweep_inflate.py --streams 65536 --cfgs "" --lz "0,0;1,0;1,100;1,400;1,1500;2,0;2,100;2,400;2,1500" > gpurun_out/r41_sweep.log 2>&1
cat gpurun_out/r41_sweep.log
