#!/bin/bash
# round 2, GPU call 13: unrolled pair checks, rank-spread staging slots, wide tiles for ClusterThrottle-heavy tables
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c13_pytest.log
tail -4 gpurun_out/c13_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c13_trace_C2.log 2>&1
timeout 120 python tools/pass_trace.py C3 > gpurun_out/c13_trace_C3.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
tail -3 gpurun_out/c13_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c13_bench.json"))
    print("pass_us %.2f frac %.3f (moved %.3f) | flush-mode %.2f us" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["roofline"]["frac_moved"], d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
    e = d["e2e"]
    print("e2e %.3g (serial %.3g, pipelined %s, separate calls %s) floor %.3g frac %.2f" % (e["value"], e["serial"]["value"], e["double_buffered"]["value"], e["separate_calls"]["value"], e["link_floor_value"], e["frac_of_link_floor"]))
    for c in d["configs"]:
        print(c.get("name"), c.get("error") or ("%.1f us frac %.3f moved %.3f value %.3g" % (c["ms_per_step"] * 1e3, c["roofline"]["frac"], c["roofline"]["frac_moved"], c["value"])))
    print(d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -30 gpurun_out/c13_trace_C2.log | head -14; grep -A12 "pass 3" gpurun_out/c13_trace_C3.log | head -30
