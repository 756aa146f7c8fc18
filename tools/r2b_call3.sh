#!/bin/bash
# round 2 (second session), GPU call 3: slimmed pass (rolled cold loops, one copy of eval_word / status / staging code), shared decide v2 without dry runs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/b3_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b3_pytest.log
tail -4 gpurun_out/b3_pytest.log
timeout 400 tools/sweep_run.sh > gpurun_out/b3_sweep_C2.log 2>&1
cat gpurun_out/b3_sweep_C2.log | cut -c1-170
KT_B200_LIB=$PWD/build/variants/libkt_b_new.so timeout 120 python tools/pass_trace.py C2 > gpurun_out/b3_trace_b_new.log 2>&1
mkdir -p build/hold && mv build/variants/libkt_c_bulkrows.so build/hold/
timeout 300 tools/sweep_run.sh --rows-scale 10 > gpurun_out/b3_sweep_x10.log 2>&1
timeout 300 tools/sweep_run.sh --config C4 > gpurun_out/b3_sweep_C4.log 2>&1
timeout 300 tools/sweep_run.sh --config C3 > gpurun_out/b3_sweep_C3.log 2>&1
cat gpurun_out/b3_sweep_x10.log gpurun_out/b3_sweep_C4.log gpurun_out/b3_sweep_C3.log | cut -c1-170
grep -A 40 "pass 3" gpurun_out/b3_trace_b_new.log | grep -v slowest | head -44
