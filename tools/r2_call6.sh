#!/bin/bash
# round 2, GPU call 6 (2 GPUs): the push exchange inside the pass -- parity, timeout behaviour, weak scaling, C3@2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c6_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c6_pytest.log
tail -15 gpurun_out/c6_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err
tail -5 gpurun_out/c6_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c6_bench_n1.json 2> gpurun_out/c6_bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/c6_bench_n1.json", "gpurun_out/c6_bench_n2.json"):
    try:
        d = json.load(open(f))
        print(f, "n_gpus", d["n_gpus"], "pass_us %.2f value %.3g" % (d["ms_per_step"] * 1e3, d["value"]), "flush-mode us %.2f" % (d["roofline"]["other_timing"]["ms_per_step"] * 1e3))
        for c in d.get("configs") or []:
            print("   ", c.get("name"), c.get("error") or ("%.1f us value %.3g" % (c["ms_per_step"] * 1e3, c["value"])))
    except Exception as e:
        print(f, "failed", e)
PY
