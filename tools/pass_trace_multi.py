"""Timeline of the fused pass on N GPUs (torchrun): where the in-kernel exchange spends its time.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/pass_trace_multi.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import kube_throttler_b200 as kt
from kube_throttler_b200 import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
snap = synth.generate("C2", seed=2 + 1000 * rank, calibrate=False)
eng = kt.Engine(snap.R, snap.L, snap.LN, device=local)
uid = [kt.Engine.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(uid[0], world, rank)
eng.upload_snapshot(snap)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
al = torch.zeros(1, device="cuda")
s = torch.cuda.Stream(); eng.set_stream(s.cuda_stream); eng.enable_trace(True)
for it in range(5):
    with torch.cuda.stream(s):
        flush.zero_(); flush.sum(); dist.all_reduce(al)
        eng.evaluate(snap.now)
    torch.cuda.synchronize()
    rows, roles = eng.trace()
    if it < 2:
        continue
    t0 = rows[:, 2].min()
    nm, nr, nf = int(roles[0]), int(roles[1]), int(roles[2])
    rec, fin, dec = rows[nm:nm + nr], rows[nm + nr:nm + nr + nf], rows[nm + nr + nf:]
    us = lambda a: (a.astype(np.float64) - float(t0)) / 1e3
    print(f"[rank {rank}] pass {it}: launches {eng.timing().launches} span {us(rows[:,3]).max():.1f} us | reconcile end max {us(rec[:,3]).max():.1f} | finalize: pre-records {np.median(us(fin[:,4])):.1f}, "
          f"own reconcile seen {np.median(us(fin[:,5])):.1f} (max {us(fin[:,5]).max():.1f}), totals of all ranks read {np.median(us(fin[:,6])):.1f} (max {us(fin[:,6]).max():.1f}), done {us(fin[:,3]).max():.1f} | "
          f"decide: words+pre staged {np.median(us(dec[:,5])):.1f}, sums seen {np.median(us(dec[:,6])):.1f} (max {us(dec[:,6]).max():.1f}), decided {np.median(us(dec[:,8])):.1f}, end max {us(dec[:,3]).max():.1f}", flush=True)
eng.close()
dist.destroy_process_group()
