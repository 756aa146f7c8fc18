#!/bin/bash
# gpurun with retries on "busy" (exit 3): tools/gpu_retry.sh TIMEOUT_S CMD...
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
