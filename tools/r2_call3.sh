#!/bin/bash
# round 2, GPU call 3: parity of the two-phase reconcile / winfo / chunked decide, traces, variant sweep, full bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c3_pytest.log
tail -4 gpurun_out/c3_pytest.log
timeout 120 python tools/pass_trace.py C2 > gpurun_out/c3_trace_C2.log 2>&1
timeout 120 python tools/pass_trace.py C2 unsorted > gpurun_out/c3_trace_C2u.log 2>&1
timeout 120 python tools/pass_trace.py C3 > gpurun_out/c3_trace_C3.log 2>&1
timeout 600 tools/sweep_run.sh > gpurun_out/c3_sweep_C2.log 2>&1
timeout 400 tools/sweep_run.sh --rows-scale 10 > gpurun_out/c3_sweep_C2x10.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c3_bench_ref.json 2> gpurun_out/c3_bench_ref.err
cat gpurun_out/c3_sweep_C2.log; cat gpurun_out/c3_sweep_C2x10.log; tail -5 gpurun_out/c3_bench.err
sed -n 1,200p gpurun_out/c3_trace_C2.log | tail -42
