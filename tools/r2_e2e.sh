#!/bin/bash
# e2e ring depth experiment: headline bench line only
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/e2e_bench.json 2> gpurun_out/e2e_bench.err
tail -3 gpurun_out/e2e_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/e2e_bench.json"))
e = d["e2e"]
print("pass_us %.2f" % (d["ms_per_step"] * 1e3), "e2e %.3g serial %.3g" % (e["value"], e["serial"]["value"]), "ring", e["double_buffered"], "link", e["host_link_gbs"], "floor %.3g frac %.2f" % (e["link_floor_value"], e["frac_of_link_floor"]))
PY
