#!/bin/bash
# round 2, GPU call 8 (8 GPUs): never-run-before N=8: parity, weak scaling, C5@8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_parity.py C3 > gpurun_out/c8_parity8.log 2>&1; echo "parity rc $?" >> gpurun_out/c8_parity8.log
grep -E "parity|MISMATCH|rc " gpurun_out/c8_parity8.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c8_bench_n8.json 2> gpurun_out/c8_bench_n8.err
tail -3 gpurun_out/c8_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c8_bench_n4.json 2> gpurun_out/c8_bench_n4.err
python - <<'PY'
import json
for f in ("gpurun_out/c8_bench_n4.json", "gpurun_out/c8_bench_n8.json"):
    try:
        d = json.load(open(f))
        print(f, "n_gpus", d["n_gpus"], "pass_us %.2f value %.3g" % (d["ms_per_step"] * 1e3, d["value"]), "flush-mode us %.2f" % (d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
        for c in d.get("configs") or []:
            print("   ", c.get("name"), c.get("error") or ("%.1f us value %.3g" % (c["ms_per_step"] * 1e3, c["value"])))
    except Exception as e:
        print(f, "failed", e)
PY
