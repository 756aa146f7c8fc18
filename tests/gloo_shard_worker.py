"""Worker of tests/test_multi_gpu.py::test_sharded_pass_with_gloo_allreduce (torchrun, gloo, CPU).

Each rank evaluates ITS row shard with the oracle (standing in for the device), all-reduces the per-throttle partials
{used[R][m], present flags, pod count} as int64 sums -- the exact buffer layout and reduction the GPUs exchange with
ncclAllReduce(int64, sum) -- and checks: reduced sums == whole-snapshot reconcile; shard pending check given the reduced
status == the whole-snapshot rows of that shard."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kube_throttler_b200 import abi, shard, synth  # noqa: E402
from oracle import ko  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for cfg, kw in (("C3", dict(m=300, n=6001, p=803)), ("C4", dict(m=400, n=5000, p=700))):
    full = synth.generate(cfg, **kw)
    mine = shard.shard_snapshot(full, rank, world)
    part = ko.columnar_evaluate(mine)  # used / used_present / used_cnt of a shard ARE its partial sums
    m, R = full.m, full.R
    buf = np.zeros((2 * R + 1, m), np.int64)  # the device layout: [used R][present R][count]
    buf[:R] = part.used
    buf[R:2 * R] = (part.used_present[None, :] >> np.arange(R, dtype=np.uint32)[:, None]) & 1
    buf[2 * R] = part.used_cnt
    t = torch.from_numpy(buf)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    want = ko.columnar_evaluate(full)
    live = ((full.thr_flags & abi.THR_RESPONSIBLE) != 0) & ((full.thr_flags & abi.THR_SELECTOR_ERROR) == 0)
    red_present = ((buf[R:2 * R] > 0).astype(np.uint32) << np.arange(R, dtype=np.uint32)[:, None]).sum(axis=0).astype(np.uint32)
    red_present |= np.where(buf[2 * R] > 0, np.uint32(abi.COUNT_BIT), np.uint32(0))
    ok &= np.array_equal(buf[:R][:, live], want.used[:, live])
    ok &= np.array_equal(buf[2 * R][live], want.used_cnt[live])
    ok &= np.array_equal(red_present[live], want.used_present[live])
    # pending rows of this shard, checked against the status every rank now agrees on
    mine.status = dict(calculated=np.ones(m, np.uint8), calc_thr=want.calc_thr, calc_present=want.calc_present, calc_cnt=want.calc_cnt,
                       used=want.used, used_present=want.used_present, used_cnt=want.used_cnt, throttled=want.throttled)
    mine.normalize()
    chk = ko.columnar_evaluate(mine, abi.EVAL_GIVEN_STATUS | abi.EVAL_SKIP_RECONCILE, words_per_row=want.words_per_row)
    lo, hi = shard.row_range(full.pending.n, rank, world)
    ok &= np.array_equal(chk.codes, want.codes[lo:hi]) and np.array_equal(chk.admit, want.admit[lo:hi])
    lo, hi = shard.row_range(full.running.n, rank, world)
    ok &= np.array_equal(part.run_bitmap, want.run_bitmap[lo:hi])
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("gloo shard parity:", "OK" if int(flag) == 1 else "FAILED")
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
