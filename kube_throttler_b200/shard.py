"""Row sharding of a snapshot across ranks: running and pending pods are split by
contiguous row range, throttles / namespaces / tables are replicated.  The only cross-rank dependency of the
pass is the per-throttle partial sums, exchanged with ONE int64 sum all-reduce (kt_comm_* -> ncclAllReduce
on the GPUs; the CPU tests use gloo for the same contract).  Pure indexing -- nothing is computed here."""
from __future__ import annotations

import copy

import numpy as np

from . import abi


def row_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous shard; the first n % world ranks get one row more."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_snapshot(snap: abi.Snapshot, rank: int, world: int) -> abi.Snapshot:
    out = copy.copy(snap)
    out.meta = dict(snap.meta, rank=rank, world=world)
    for name in ("running", "pending"):
        pods = getattr(snap, name)
        lo, hi = row_range(pods.n, rank, world)
        setattr(out, name, pods.rows(slice(lo, hi)))
    return out.normalize()


# per-throttle results are identical on every rank after the all-reduce; per-pod results are row-sharded
PER_POD = ("run_bitmap", "pend_bitmap", "codes", "admit")


def concat_results(parts: list[abi.PassResult]) -> abi.PassResult:
    """Stack the per-rank row shards back into one result (rank order == row order)."""
    out = copy.copy(parts[0])
    for f in PER_POD:
        setattr(out, f, np.concatenate([getattr(p, f) for p in parts], axis=0))
    return out
