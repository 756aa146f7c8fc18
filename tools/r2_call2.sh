#!/bin/bash
# round 2, GPU call 2: parity of the restructured pass (finalize off the critical path), launch floor, trace, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c2_pytest.log
tail -4 gpurun_out/c2_pytest.log
timeout 120 tools/micro/launch_floor > gpurun_out/c2_launch_floor.txt 2>&1
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c2_trace_C2.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 --rows-scale 10 > gpurun_out/c2_bench_x10.json 2> gpurun_out/c2_bench_x10.err
KT_ROWS_SCALE=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/c2_prof_x10 python tools/ncu_target.py C2 4 fused > gpurun_out/c2_ncu_x10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/c2_prof_C2 python tools/ncu_target.py C2 4 fused > gpurun_out/c2_ncu_C2.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/c2_bench.json", "gpurun_out/c2_bench_x10.json"):
    try:
        d = json.load(open(f)); print(f, "pass_us %.2f" % (d["ms_per_step"] * 1e3), "frac %.3f" % d["roofline"]["frac"], "e2e %.3g" % d["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -12 gpurun_out/c2_trace_C2.log
