#!/bin/bash
# round 2, GPU call 14: first two words evaluated with interleaved gathers (A/B), loop pair checks restored
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c14_pytest.log
tail -4 gpurun_out/c14_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c14_trace_C2.log 2>&1
timeout 300 tools/sweep_run.sh > gpurun_out/c14_sweep_C2.log 2>&1
timeout 300 tools/sweep_run.sh --config C4 > gpurun_out/c14_sweep_C4.log 2>&1
timeout 300 tools/sweep_run.sh --rows-scale 10 > gpurun_out/c14_sweep_x10.log 2>&1
cat gpurun_out/c14_sweep_C2.log gpurun_out/c14_sweep_C4.log gpurun_out/c14_sweep_x10.log | cut -c1-150
tail -30 gpurun_out/c14_trace_C2.log | head -24
