#!/bin/bash
# compute-sanitizer over the small fixtures of tools/sanitize_case.py — the analogue of the reference's valgrind job
# (/root/reference/.github/workflows/rust.yml:79-83). ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'tools/sanitize.sh memcheck'      then, in another call,   'tools/sanitize.sh racecheck'
# The summary lands in gpurun_out/sanitize_<tool>.log; copy it to profiles/ for the record.
# Round 2: this pool answered "compute-sanitizer is closed on this pool and stays closed" (exit code 86), so no summary could be
# committed; the fixtures still run plainly (first step below) and the CPU emulator (tests/cusim: guard bytes, randomised
# scheduling of the kernel sources' threads) carries the memory / race checking.
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/sanitize_${tool}.log 2>&1
rc=$?
echo "compute-sanitizer --tool $tool: exit code $rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case ok|========= (Invalid|Race|Error|Hazard)" gpurun_out/sanitize_${tool}.log | head -40
exit $rc
