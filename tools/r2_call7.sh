#!/bin/bash
# round 2, GPU call 7 (2 GPUs): tagged-word exchange -- parity, timeout, trace, weak scaling
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c7_pytest.log
tail -6 gpurun_out/c7_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/pass_trace_multi.py 2>&1 | grep "rank" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err
tail -3 gpurun_out/c7_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/c7_bench_n2.json",):
    try:
        d = json.load(open(f))
        print(f, "n_gpus", d["n_gpus"], "pass_us %.2f value %.3g" % (d["ms_per_step"] * 1e3, d["value"]), "flush-mode us %.2f" % (d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
    except Exception as e:
        print(f, "failed", e)
PY
