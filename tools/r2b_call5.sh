#!/bin/bash
# round 2 (second session), GPU call 5 (2 GPUs): A/B of the shared decide tile's acquire (one fence vs three acquire loads) on one GPU, then the
# two-GPU parity tests, C2-shaped parity, and the N=2 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  local so=$1; shift
  CUDA_VISIBLE_DEVICES=0 KT_B200_LIB=$PWD/build/variants/libkt_$so.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 1 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$so', '$*', 'pass_us %.2f' % (d['ms_per_step']*1e3), 'flush_us %.2f' % (d['roofline']['other_timing']['ms_per_step']*1e3))"
}
for r in 1 2; do for v in a_old b_new c_acq; do run $v; done; done > gpurun_out/b5_sweep.log 2>&1
for v in a_old b_new c_acq; do run $v --config C3; done >> gpurun_out/b5_sweep.log 2>&1
cat gpurun_out/b5_sweep.log
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/b5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b5_pytest.log
tail -5 gpurun_out/b5_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/multi_gpu_parity.py C2 1000 100000 10000 > gpurun_out/b5_parity_C2.log 2>&1
tail -3 gpurun_out/b5_parity_C2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/b5_bench_n2.err
tail -3 gpurun_out/b5_bench_n2.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_n2.json"))
    print("n_gpus", d["n_gpus"], "pass_us %.2f value %.3g" % (d["ms_per_step"] * 1e3, d["value"]), "e2e %.3g" % d["e2e"]["value"])
    for c in d.get("configs", []):
        print(c["name"], "%.1f us" % (c["ms_per_step"] * 1e3), "value %.3g" % c["value"])
except Exception as e:
    print("failed", e)
PY
