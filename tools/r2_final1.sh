#!/bin/bash
# round 2, final single-GPU evidence of the committed build: GPU suite, bench line, reference arm, in-kernel trace, ncu captures + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/f1_pytest.log
tail -4 gpurun_out/f1_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/r2_trace_C2.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/f1_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/f1_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/r2_prof_C2 -f python tools/ncu_target.py C2 4 fused > gpurun_out/f1_ncu_C2.log 2>&1
KT_ROWS_SCALE=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/r2_prof_C2x10 -f python tools/ncu_target.py C2 4 fused > gpurun_out/f1_ncu_x10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/r2_prof_C3 -f python tools/ncu_target.py C3 4 fused > gpurun_out/f1_ncu_C3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 2 > gpurun_out/f1_launches.log 2>&1
tail -3 gpurun_out/f1_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench.json"))
    print("pass_us %.2f frac %.3f (moved %.3f) | flush-mode %.2f us" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["roofline"]["frac_moved"], d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
    e = d["e2e"]
    print("e2e %.3g (serial %.3g, pipelined %s, separate calls %s) floor %.3g frac %.2f link %s" % (e["value"], e["serial"]["value"], e["double_buffered"]["value"], e["separate_calls"]["value"], e["link_floor_value"], e["frac_of_link_floor"], e["host_link_gbs"]))
    for c in d["configs"]:
        print(c.get("name"), c.get("error") or ("%.1f us frac %.3f moved %.3f value %.3g" % (c["ms_per_step"] * 1e3, c["roofline"]["frac"], c["roofline"]["frac_moved"], c["value"])))
    print(json.dumps(d["e2e_plugin"])[:2500])
    print(d["cpu_baseline"])
    r = json.load(open("gpurun_out/r2_bench_reference.json"))
    print("reference arm", r["value"], r["config"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -30 gpurun_out/r2_trace_C2.txt | head -22
