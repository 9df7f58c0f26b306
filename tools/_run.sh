for v in 7 3; do
CZ_MATCH_V=$v python bench.py --workload deflate --mib 1024 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r63_deflate_v$v.json 2> gpurun_out/r63_deflate_v$v.err
echo "v$v: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r63_deflate_v$v.json | head -1)"; tail -1 gpurun_out/r63_deflate_v$v.err | cut -c1-200
done
