#!/bin/bash
# Round-end evidence run on a B200 box (under gpurun): the bench line, `ncu --set full` captures of the product kernels of every
# workload, and the launch list of the whole bench. The reports stay on the box (gpurun brings back at most 64 MiB): what comes
# back in gpurun_out/ under the prefix $1 are their summaries (tools/ncu_summary.py), per-source-line roll-ups of the dominant
# kernels (tools/ncu_src.py) and the launch list.
N=${1:-cap}
P=gpurun_out/$N
T=/tmp/$N
python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo bench_rc=$?
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'inflate_tok_kernel|inflate_lz_kernel' -s 2 -c 2 -o ${T}_inflate python bench.py --workload inflate --steps 1 --warmup 1 --no-cpu --no-e2e > ${P}_ncu_inflate.log 2>&1; echo ncu_inflate_rc=$?
python tools/ncu_summary.py ${T}_inflate.ncu-rep ${P}_inflate_ncu.md --alg-bytes 6322910765
python tools/ncu_src.py ${T}_inflate.ncu-rep inflate_tok 30 > ${P}_inflate_tok_lines.txt 2>&1
python tools/ncu_src.py ${T}_inflate.ncu-rep inflate_lz 30 > ${P}_inflate_lz_lines.txt 2>&1
$NCU -k regex:'deflate_' -s 12 -c 14 -o ${T}_deflate python bench.py --workload deflate --mib 1024 --sub-steps 1 --no-cpu --no-e2e > ${P}_ncu_deflate.log 2>&1; echo ncu_deflate_rc=$?
python tools/ncu_summary.py ${T}_deflate.ncu-rep ${P}_deflate_ncu.md
python tools/ncu_src.py ${T}_deflate.ncu-rep match_sweep 30 > ${P}_deflate_match_lines.txt 2>&1
$NCU -k regex:'inflate_candidates|inflate_tok_kernel|inflate_lz16|inflate_tail_markers|inflate_window|inflate_resolve|inflate_run_check' -s 8 -c 8 -o ${T}_runs python bench.py --workload gzip --sub-steps 1 > ${P}_ncu_runs.log 2>&1; echo ncu_runs_rc=$?
python tools/ncu_summary.py ${T}_runs.ncu-rep ${P}_runs_ncu.md
python tools/ncu_src.py ${T}_runs.ncu-rep candidates 25 > ${P}_runs_candidates_lines.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 1 --sub-steps 1 --no-cpu > ${P}_ncu_launches.log 2>&1; echo ncu_launches_rc=$?
ls -la ${T}_*.ncu-rep
du -sh gpurun_out
