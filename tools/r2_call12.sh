#!/bin/bash
# round 2, GPU call 12: resident pass (match CTAs stay on as decide tiles), one-block uploads, tile 128 vs 256 per shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c12_pytest.log
tail -4 gpurun_out/c12_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c12_trace_C2.log 2>&1
for cfg in C2 C3 C4 C5; do
  timeout 400 tools/sweep_run.sh --config $cfg > gpurun_out/c12_sweep_$cfg.log 2>&1
  echo "== $cfg"; grep "round 2" gpurun_out/c12_sweep_$cfg.log | cut -c1-140
done
timeout 300 tools/sweep_run.sh --rows-scale 10 > gpurun_out/c12_sweep_C2x10.log 2>&1
echo "== C2x10"; grep "round 2" gpurun_out/c12_sweep_C2x10.log | cut -c1-140
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
tail -3 gpurun_out/c12_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c12_bench.json"))
    print("pass_us %.2f frac %.3f (moved %.3f) | flush-mode %.2f us" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["roofline"]["frac_moved"], d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
    e = d["e2e"]
    print("e2e %.3g (serial %.3g, pipelined %s, separate calls %s) floor %.3g frac %.2f" % (e["value"], e["serial"]["value"], e["double_buffered"]["value"], e["separate_calls"]["value"], e["link_floor_value"], e["frac_of_link_floor"]))
    for c in d["configs"]:
        print(c.get("name"), c.get("error") or ("%.1f us frac %.3f moved %.3f value %.3g" % (c["ms_per_step"] * 1e3, c["roofline"]["frac"], c["roofline"]["frac_moved"], c["value"])))
except Exception as e:
    print("bench parse failed", e)
PY
tail -30 gpurun_out/c12_trace_C2.log
