python bench.py --workload deflate --mib 1024 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r60_deflate.json 2> gpurun_out/r60_deflate.err
echo "$(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r60_deflate.json | head -1)"; tail -1 gpurun_out/r60_deflate.err | cut -c1-200
