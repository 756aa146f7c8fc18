#!/bin/bash
# round 2, GPU call 11: prep inside match tiles; host reconcile on the device diff; tile 128 vs 256 on every shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c11_pytest.log
tail -4 gpurun_out/c11_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c11_trace_C2.log 2>&1
for cfg in C2 C3 C4 C5; do
  timeout 400 tools/sweep_run.sh --config $cfg > gpurun_out/c11_sweep_$cfg.log 2>&1
  echo "== $cfg"; grep "round 2" gpurun_out/c11_sweep_$cfg.log | cut -c1-120
done
timeout 300 tools/sweep_run.sh --rows-scale 10 > gpurun_out/c11_sweep_C2x10.log 2>&1
echo "== C2x10"; grep "round 2" gpurun_out/c11_sweep_C2x10.log | cut -c1-120
tail -24 gpurun_out/c11_trace_C2.log
