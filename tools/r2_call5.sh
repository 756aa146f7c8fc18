#!/bin/bash
# round 2, GPU call 5: maintained bitmaps (no zero-fill), multi-word decide rounds, scattered-warp paths, step API, queue bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c5_pytest.log
tail -4 gpurun_out/c5_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c5_trace_C2.log 2>&1
timeout 120 python tools/pass_trace.py C2 unsorted > gpurun_out/c5_trace_C2u.log 2>&1
timeout 120 python tools/pass_trace.py C3 > gpurun_out/c5_trace_C3.log 2>&1
timeout 400 tools/sweep_run.sh > gpurun_out/c5_sweep_C2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
cat gpurun_out/c5_sweep_C2.log; tail -5 gpurun_out/c5_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c5_bench.json"))
    print("pass_us %.2f frac %.3f | flush-mode %.2f us" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
    e = d["e2e"]
    print("e2e %.3g (serial %.3g, pipelined %s, separate calls %s) floor %.3g frac %.2f" % (e["value"], e["serial"]["value"], e["double_buffered"]["value"], e["separate_calls"]["value"], e["link_floor_value"], e["frac_of_link_floor"]))
    for c in d["configs"]:
        print(c.get("name"), c.get("error") or ("%.1f us frac %.3f value %.3g" % (c["ms_per_step"] * 1e3, c["roofline"]["frac"], c["value"])))
    print(json.dumps(d["e2e_plugin"])[:3500])
    print(d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -34 gpurun_out/c5_trace_C2.log
