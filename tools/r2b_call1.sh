#!/bin/bash
# round 2 (second session), GPU call 1: shared decide tiles + bulk-copy staging -- parity of the default build, in-kernel trace, A/B sweep at C2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/b1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b1_pytest.log
tail -6 gpurun_out/b1_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/b1_trace_C2.log 2>&1
timeout 420 tools/sweep_run.sh > gpurun_out/b1_sweep_C2.log 2>&1
cat gpurun_out/b1_sweep_C2.log | cut -c1-170
tail -42 gpurun_out/b1_trace_C2.log | head -40
