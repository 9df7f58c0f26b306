for mb in 0 256 512 1024; do
CZ_INFLATE_FAST_MB=$mb python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r65_bench_f$mb.json 2> gpurun_out/r65_bench_f$mb.err
echo "fast $mb: $(grep -o '"e2e": {[^}]*}' gpurun_out/r65_bench_f$mb.json)"; tail -1 gpurun_out/r65_bench_f$mb.err | cut -c1-200
done
