#!/bin/bash
# compute-sanitizer over every device path of the round-2 build (tools/sanitize_target.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/san_plain.log 2>&1; echo "plain rc $?"; tail -2 gpurun_out/san_plain.log
KT_SANITIZE_BIG=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc $?"
tail -3 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc $?"
tail -3 gpurun_out/san_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/san_synccheck.log 2>&1; echo "synccheck rc $?"
tail -3 gpurun_out/san_synccheck.log
