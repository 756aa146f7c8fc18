#!/bin/bash
# round 2 (second session), GPU call 2: primitive latencies, shared decide v2 (merged polls, pairs dealt to four lanes, dry runs): parity, traces, A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/micro/prims > gpurun_out/b2_prims.log 2>&1
cat gpurun_out/b2_prims.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/b2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b2_pytest.log
tail -6 gpurun_out/b2_pytest.log
for v in b_new c_dry0 d_dry2; do
  KT_B200_LIB=$PWD/build/variants/libkt_$v.so timeout 120 python tools/pass_trace.py C2 > gpurun_out/b2_trace_$v.log 2>&1
done
timeout 400 tools/sweep_run.sh > gpurun_out/b2_sweep_C2.log 2>&1
cat gpurun_out/b2_sweep_C2.log | cut -c1-170
mkdir -p build/hold && mv build/variants/libkt_c_dry0.so build/variants/libkt_d_dry2.so build/variants/libkt_f_res1.so build/hold/
timeout 300 tools/sweep_run.sh --rows-scale 10 > gpurun_out/b2_sweep_x10.log 2>&1
timeout 300 tools/sweep_run.sh --config C4 > gpurun_out/b2_sweep_C4.log 2>&1
cat gpurun_out/b2_sweep_x10.log gpurun_out/b2_sweep_C4.log | cut -c1-170
for v in b_new c_dry0 d_dry2; do echo "== trace $v"; grep -A 40 "pass 3" gpurun_out/b2_trace_$v.log | grep -v slowest | head -44; done
