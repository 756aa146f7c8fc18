#!/bin/bash
# round 2 (second session), GPU call 7: compute-sanitizer (memcheck, racecheck) over the resident pass with the shared second phase, small cases
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KT_SANITIZE_SMALL=1 timeout 50 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/b7_memcheck.log 2>&1; echo "memcheck rc $?"
tail -3 gpurun_out/b7_memcheck.log
KT_SANITIZE_SMALL=1 timeout 50 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/b7_racecheck.log 2>&1; echo "racecheck rc $?"
tail -3 gpurun_out/b7_racecheck.log
