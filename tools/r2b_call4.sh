#!/bin/bash
# round 2 (second session), GPU call 4: candidate final build (shared second phase in resident passes; reconcile / match tiles as before):
# whole GPU suite, A/B against the build at the start of the session on four shapes, trace, bench line, ncu capture + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/b4_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b4_pytest.log
tail -4 gpurun_out/b4_pytest.log
run() {  # run LIB ARGS...: one bench line, reduced
  local so=$1; shift
  KT_B200_LIB=$PWD/build/variants/libkt_$so.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 1 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$so', '$*', 'pass_us %.2f' % (d['ms_per_step']*1e3), 'flush_us %.2f' % (d['roofline']['other_timing']['ms_per_step']*1e3), 'e2e %.3g' % d['e2e']['value'])"
}
for r in 1 2; do for v in a_old b_new d_res1; do run $v; done; done > gpurun_out/b4_sweep.log 2>&1
for v in a_old b_new; do run $v --config C3; run $v --rows-scale 10; run $v --config C4; done >> gpurun_out/b4_sweep.log 2>&1
cat gpurun_out/b4_sweep.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/b4_trace_C2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/b4_bench.err
tail -2 gpurun_out/b4_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 2 -c 1 -o gpurun_out/r2b_prof_C2 -f python tools/ncu_target.py C2 4 fused > gpurun_out/b4_ncu_C2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 2 > gpurun_out/b4_launches.log 2>&1
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/b4_ref.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2b_bench.json"))
    print("pass_us %.2f frac %.3f (moved %.3f) | flush-mode %.2f us" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["roofline"]["frac_moved"], d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
    e = d["e2e"]
    print("e2e %.3g (serial %.3g) floor %.3g frac %.2f" % (e["value"], e["serial"]["value"], e["link_floor_value"], e["frac_of_link_floor"]))
    for c in d["configs"]:
        print(c.get("name"), c.get("error") or ("%.1f us frac %.3f value %.3g" % (c["ms_per_step"] * 1e3, c["roofline"]["frac"], c["value"])))
    print(d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
grep -A 40 "pass 3" gpurun_out/b4_trace_C2.log | grep -v slowest | head -40
