#!/bin/bash
# round 2, GPU call 1: parity of the merged tree, variant sweep, in-kernel trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c1_pytest.log
timeout 600 tools/sweep_run.sh > gpurun_out/c1_sweep_C2.log 2>&1
timeout 300 tools/sweep_run.sh --rows-scale 10 > gpurun_out/c1_sweep_C2x10.log 2>&1
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c1_trace_C2.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_sweep_C2.log; cat gpurun_out/c1_trace_C2.log | tail -8
