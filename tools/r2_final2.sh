#!/bin/bash
# round 2, final build on N GPUs (N = $1, default 2): multi-GPU parity tests + the bench line with the BASELINE shape of that N
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$N" = 2 ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/f2_pytest.log
  tail -4 gpurun_out/f2_pytest.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/f2_bench_n$N.err
tail -3 gpurun_out/f2_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n$N.json"))
    print("n_gpus", d["n_gpus"], "pass_us %.2f value %.3g" % (d["ms_per_step"] * 1e3, d["value"]), "e2e %.3g" % d["e2e"]["value"])
    for c in d.get("configs", []):
        print(c["name"], "%.1f us" % (c["ms_per_step"] * 1e3), "value %.3g" % c["value"])
except Exception as e:
    print("failed", e)
PY
