python tools/sweep_inflate.py --streams 65536 --cfgs=-2,14 --steps 5 > gpurun_out/r84_sweep.log 2>&1; tail -1 gpurun_out/r84_sweep.log | cut -c1-120
CZ_NO_L1_PREF=1 python tools/sweep_inflate.py --streams 65536 --cfgs=-2,14 --steps 5 > gpurun_out/r84_sweep_no.log 2>&1; tail -1 gpurun_out/r84_sweep_no.log | cut -c1-120
python bench.py --workload deflate --mib 1024 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r84_d.json 2>/dev/null; echo "deflate l1pref: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r84_d.json | head -1)"
CZ_NO_L1_PREF=1 python bench.py --workload deflate --mib 1024 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r84_d_no.json 2>/dev/null; echo "deflate no pref: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r84_d_no.json | head -1)"
