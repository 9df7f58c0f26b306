for cfg in "12 16" "13 18" "10 14" "7 10" "14 28"; do
set -- $cfg
CZ_CHAIN_PER_SM=$1 CZ_PARSE_PER_SM=$2 python bench.py --workload deflate --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/r66_deflate.json 2> gpurun_out/r66_deflate.err
echo "chain/SM $1 parse/SM $2: $(grep -o 'ms_per_step": [0-9.]*' gpurun_out/r66_deflate.json | head -1)"; tail -1 gpurun_out/r66_deflate.err | cut -c1-200
done
