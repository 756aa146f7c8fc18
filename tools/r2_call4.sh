#!/bin/bash
# round 2, GPU call 4: CTA-level decide staging, PDL launch, resident queue on the GPU, device status diff
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c4_pytest.log
tail -4 gpurun_out/c4_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c4_trace_C2.log 2>&1
timeout 120 python tools/pass_trace.py C3 > gpurun_out/c4_trace_C3.log 2>&1
timeout 400 tools/sweep_run.sh > gpurun_out/c4_sweep_C2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
cat gpurun_out/c4_sweep_C2.log; tail -5 gpurun_out/c4_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c4_bench.json"))
    print("pass_us %.2f frac %.3f e2e %.3g floor %.3g" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["link_floor_value"]))
    for c in d["configs"]:
        print(c.get("name"), c.get("error") or ("%.1f us frac %.3f value %.3g" % (c["ms_per_step"] * 1e3, c["roofline"]["frac"], c["value"])))
    print(json.dumps(d["e2e_plugin"])[:3000])
    print(d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -32 gpurun_out/c4_trace_C2.log
