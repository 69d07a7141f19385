#!/bin/bash
# compute-sanitizer over the step kernels (tools/sanitize_cases.py); logs under gpurun_out/sanitizer/.
# memcheck on the full case list, racecheck + initcheck-free synccheck on the quick list (racecheck is ~50x slower).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sanitizer
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
timeout 900 $CS --tool memcheck --error-exitcode 7 --log-file gpurun_out/sanitizer/memcheck.log \
    python tools/sanitize_cases.py > gpurun_out/sanitizer/memcheck.out 2>&1
echo "memcheck exit $?" | tee gpurun_out/sanitizer/summary.txt
LBM_SPECULATIVE=1 timeout 900 $CS --tool memcheck --error-exitcode 7 --log-file gpurun_out/sanitizer/memcheck_speculative.log \
    python tools/sanitize_cases.py quick > gpurun_out/sanitizer/memcheck_speculative.out 2>&1
echo "memcheck (speculative pull forced) exit $?" | tee -a gpurun_out/sanitizer/summary.txt
timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 7 --log-file gpurun_out/sanitizer/racecheck.log \
    python tools/sanitize_cases.py quick > gpurun_out/sanitizer/racecheck.out 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/sanitizer/summary.txt
tail -3 gpurun_out/sanitizer/memcheck.log gpurun_out/sanitizer/memcheck_speculative.log gpurun_out/sanitizer/racecheck.log | tee -a gpurun_out/sanitizer/summary.txt
