"""A rank that never shows up must not wedge the others' GPUs: rank 0 runs a pass alone; its in-kernel wait for rank 1 gives up
after two seconds, the kernel ends, and the next getter reports KT_ERR_STATE.  Afterwards both ranks run a normal pass.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_timeout.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import kube_throttler_b200 as kt
from kube_throttler_b200 import abi, shard, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
full = synth.generate("C3", m=300, n=6000, p=800)
mine = shard.shard_snapshot(full, rank, world)
eng = kt.Engine(full.R, full.L, full.LN, device=local)
uid = [kt.Engine.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(uid[0], world, rank)
eng.upload_snapshot(mine)
eng.evaluate(full.now)   # pass 1, everybody: maps the exchange windows
first = eng.download()
dist.barrier()
ok = True
if rank == 0:
    t0 = time.time()
    eng.evaluate(full.now)  # pass 2, rank 0 ALONE
    try:
        eng.download()
        ok = False
        print("rank 0: the lonely pass did not report a timeout")
    except kt.KtError as e:
        dt = time.time() - t0
        ok = e.code == abi.ERR_STATE and 1.5 < dt < 10
        print(f"rank 0: lonely pass -> {e} after {dt:.1f}s: {'OK' if ok else 'UNEXPECTED'}")
dist.barrier()
if rank != 0:
    eng.evaluate(full.now)  # rank 1 catches up with its pass 2 (rank 0's epoch 2 is already published)
    eng.download()
dist.barrier()
eng.evaluate(full.now)      # pass 3, everybody again: results as in pass 1
again = eng.download()
same = all(np.array_equal(getattr(first, f), getattr(again, f)) for f in ("used", "used_cnt", "codes", "admit"))
flag = torch.tensor([1 if (ok and same) else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("multi-gpu timeout:", "OK" if int(flag.item()) == 1 else f"FAILED (ok={ok}, same={same})")
eng.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
