#!/bin/bash
# round 2, GPU call 10: prep tiles first + batched, stage chunk 4; plugin bench with reserve-invalidated queue pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c10_pytest.log
tail -4 gpurun_out/c10_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c10_trace_C2.log 2>&1
timeout 300 tools/sweep_run.sh > gpurun_out/c10_sweep_C2.log 2>&1
timeout 300 python tools/plugin_bench.py > gpurun_out/c10_plugin.json 2> gpurun_out/c10_plugin.err
cat gpurun_out/c10_sweep_C2.log; tail -3 gpurun_out/c10_plugin.err
python - <<'PY'
import json
try:
    p = json.load(open("gpurun_out/c10_plugin.json"))
    print(json.dumps(p)[:3000])
except Exception as e:
    print("plugin parse failed", e)
PY
tail -34 gpurun_out/c10_trace_C2.log
