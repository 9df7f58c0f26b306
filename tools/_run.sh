for mb in 128 320; do
CZ_TRACE=1 CZ_INFLATE_FAST_MB=$mb python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r78_trace_$mb.json 2> gpurun_out/r78_trace_$mb.err
echo "== fast $mb"; grep "\[cz\]" gpurun_out/r78_trace_$mb.err | tail -19 | cut -c5-130
grep -o '"e2e": {[^}]*}' gpurun_out/r78_trace_$mb.json | grep -o "ms_per_step.*"
done
