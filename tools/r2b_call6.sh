#!/bin/bash
# round 2 (second session), GPU call 6: final build -- whole GPU suite, bench line, reference arm, in-kernel trace, ncu captures (C2, C3) + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/b6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b6_pytest.log
tail -4 gpurun_out/b6_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/b6_bench.err
tail -2 gpurun_out/b6_bench.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/r2b_prof_C2 -f python tools/ncu_target.py C2 4 fused > gpurun_out/b6_ncu_C2.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 2 > gpurun_out/b6_launches.log 2>&1
timeout 60 python tools/pass_trace.py C2 > gpurun_out/r2b_trace_C2.txt 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/r2b_prof_C3 -f python tools/ncu_target.py C3 4 fused > gpurun_out/b6_ncu_C3.log 2>&1
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/b6_ref.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2b_bench.json"))
    print("pass_us %.2f frac %.3f (moved %.3f) | flush-mode %.2f us" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["roofline"]["frac_moved"], d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
    e = d["e2e"]
    print("e2e %.3g (serial %.3g) floor %.3g frac %.2f" % (e["value"], e["serial"]["value"], e["link_floor_value"], e["frac_of_link_floor"]))
    for c in d["configs"]:
        print(c.get("name"), c.get("error") or ("%.1f us frac %.3f value %.3g" % (c["ms_per_step"] * 1e3, c["roofline"]["frac"], c["value"])))
    print(d["cpu_baseline"]["value"], d.get("clocks"))
    r = json.load(open("gpurun_out/r2b_bench_reference.json"))
    print("reference arm", r["value"])
except Exception as e:
    print("bench parse failed", e)
PY
