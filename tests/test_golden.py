"""Committed golden vectors (tests/golden/, made by tests/golden/make_golden.py from the oracle):
  CPU: the generator still produces the same inputs and the oracle the same outputs (drift guard);
  GPU: the CUDA path matches the committed vectors bit for bit, independently of the live oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

from kube_throttler_b200 import abi, synth

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = json.load(open(os.path.join(HERE, "index.json")))
FIELDS = ("admit", "codes", "pend_bitmap", "used", "used_present", "used_cnt", "throttled", "calc_thr", "calc_present", "calc_cnt", "override_active")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    meta = INDEX[name]
    snap = synth.generate(meta["config"], **meta["kwargs"])
    inputs = digest(np.concatenate([snap.running.labels.ravel(), snap.running.req.ravel(), snap.pending.labels.ravel(), snap.thr.ravel(),
                                    snap.req_vals.astype(np.int64)]))
    assert inputs == meta["inputs_sha256"], "the synthetic generator no longer reproduces the golden inputs"
    return meta, snap, np.load(os.path.join(HERE, name + ".npz"))


def compare(snap, got, gold):
    live = ((snap.thr_flags & abi.THR_RESPONSIBLE) != 0) & ((snap.thr_flags & abi.THR_SELECTOR_ERROR) == 0)
    for f in FIELDS:
        a, b = getattr(got, f), gold[f]
        if f in ("used", "used_present", "used_cnt", "throttled"):  # defined for the throttles this instance reconciles
            a, b = (a[:, live], b[:, live]) if a.ndim == 2 else (a[live], b[live])
        np.testing.assert_array_equal(a, b, err_msg=f)
    assert digest(got.run_bitmap) == bytes(gold["run_bitmap_sha256"]).hex(), "running match bitmap"


@pytest.mark.parametrize("name", sorted(INDEX))
def test_oracle_reproduces_golden(oracle, name):
    meta, snap, gold = load(name)
    compare(snap, oracle.columnar_evaluate(snap, meta["flags"], words_per_row=meta["words_per_row"]), gold)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(INDEX))
def test_cuda_path_matches_golden(kt, name):
    meta, snap, gold = load(name)
    got = kt.evaluate_snapshot(snap, meta["flags"])
    assert got.words_per_row == meta["words_per_row"]
    compare(snap, got, gold)
