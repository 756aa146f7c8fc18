"""Pins the COLUMNAR oracle (plain loops over the engine's int64 columns) against the object-level
oracle (string maps, Quantities, per-call selector construction) and against the generator's numpy
matcher: three independent implementations must agree on every output.  CPU only."""
import numpy as np
import pytest

from kube_throttler_b200 import abi, synth


def live(snap):
    return ((snap.thr_flags & abi.THR_RESPONSIBLE) != 0) & ((snap.thr_flags & abi.THR_SELECTOR_ERROR) == 0)


def compare(snap, col, obj, reconcile=True):
    np.testing.assert_array_equal(col.pend_bitmap, obj.pend_bitmap, err_msg="pending bitmap")
    np.testing.assert_array_equal(col.codes, obj.codes, err_msg="codes")
    np.testing.assert_array_equal(col.admit, obj.admit, err_msg="admit")
    if reconcile:
        lv = live(snap)
        for f in ("used", "calc_thr"):
            np.testing.assert_array_equal(getattr(col, f)[:, lv], getattr(obj, f)[:, lv], err_msg=f)
        for f in ("used_present", "used_cnt", "throttled", "calc_present", "calc_cnt"):
            np.testing.assert_array_equal(getattr(col, f)[lv], getattr(obj, f)[lv], err_msg=f)


@pytest.mark.parametrize("kw", [
    dict(config="C2", m=64, n=2000, p=300),
    dict(config="C2", m=200, n=5000, p=500, sort_by_namespace=False),
    dict(config="C3", m=200, n=4000, p=600),
    dict(config="C4", m=300, n=4000, p=600),
    dict(config="C3", m=150, n=3000, p=400, seed=11, override_frac=0.3),
])
@pytest.mark.parametrize("flags", [abi.EVAL_FRESH_STATUS, abi.EVAL_ON_EQUAL])
def test_columnar_vs_object(oracle, kw, flags):
    kw = dict(kw)
    snap = synth.generate(kw.pop("config"), **kw)
    col = oracle.columnar_evaluate(snap, flags)
    obj, tm = oracle.object_evaluate(snap, flags, threads=4)
    assert not tm["had_error"]
    compare(snap, col, obj)
    used, present, cnt = snap.meta["true_used"]  # numpy matcher
    lv = live(snap)
    np.testing.assert_array_equal(col.used[:, lv], used[:, lv])
    np.testing.assert_array_equal(col.used_present[lv], present[lv])
    np.testing.assert_array_equal(col.used_cnt[lv], cnt[lv])
    codes = col.code_matrix(snap.m)
    assert (codes == abi.CHECK_ACTIVE).any() and (codes == abi.CHECK_INSUFFICIENT).any()


def test_given_status_columnar_vs_object(oracle):
    snap = synth.generate("C4", m=300, n=4000, p=600)
    fresh = oracle.columnar_evaluate(snap)
    rng = np.random.default_rng(5)
    m = snap.m
    st = dict(calculated=(rng.random(m) < 0.7).astype(np.uint8), calc_thr=fresh.calc_thr.copy(), calc_present=fresh.calc_present.copy(),
              calc_cnt=fresh.calc_cnt.copy(), used=(fresh.used * 0.7).astype(np.int64), used_present=fresh.used_present.copy(),
              used_cnt=fresh.used_cnt.copy(), throttled=np.where(rng.random(m) < 0.3, 0, fresh.throttled).astype(np.uint32))
    snap.status = st
    snap.normalize()
    for flags in (abi.EVAL_GIVEN_STATUS, abi.EVAL_GIVEN_STATUS | abi.EVAL_ON_EQUAL):
        col = oracle.columnar_evaluate(snap, flags)
        obj, tm = oracle.object_evaluate(snap, flags, threads=2)
        compare(snap, col, obj, reconcile=False)


def test_c1_literal(oracle):
    """BASELINE config 1 (example/throttle.yaml shape): 10 x 10m running vs cpu=200m; pending 100m/101m/300m."""
    snap = synth.generate("C1")
    col = oracle.columnar_evaluate(snap)
    assert col.used[0, 0] == 100 and col.used_cnt[0] == 10 and col.throttled[0] == 0
    assert list(col.code_matrix(1)[:, 0]) == [abi.CHECK_NOT_THROTTLED, abi.CHECK_INSUFFICIENT, abi.CHECK_POD_REQUESTS_EXCEEDS_THRESHOLD]
    assert list(col.admit) == [1, 0, 0]
    obj, _ = oracle.object_evaluate(snap)
    compare(snap, col, obj)


def test_selector_error_and_invalid_ns_term(oracle):
    """KT_THR_SELECTOR_ERROR throttles never match on the columnar side; the object side raises (plugin.go:154)."""
    snap = synth.generate("C3", m=60, n=1000, p=100)
    snap.thr_flags[3] |= abi.THR_SELECTOR_ERROR
    cl = np.nonzero(snap.kind == abi.KIND_CLUSTERTHROTTLE)[0]
    snap.term_flags[snap.term_off[cl[0]]] |= abi.TERM_NS_INVALID
    snap.normalize()
    col = oracle.columnar_evaluate(snap)
    assert col.match_matrix("pending", snap.m)[:, 3].sum() == 0
    obj, tm = oracle.object_evaluate(snap)
    ns3 = snap.thr_ns[3]
    bad = snap.pending.ns_id == ns3  # pods whose namespace holds the broken Throttle get framework.Error
    assert tm["had_error"] == bool(bad.any())
    ok = ~bad
    np.testing.assert_array_equal(col.codes[ok], obj.codes[ok])


def test_shard_union_equals_whole(oracle):
    """Row sharding: summing per-shard used and concatenating per-shard checks == one pass."""
    snap = synth.generate("C3", m=200, n=4000, p=600)
    whole = oracle.columnar_evaluate(snap)
    G = 4
    parts = [snap.shard(r, G) for r in range(G)]
    used = sum(oracle.columnar_evaluate(p, abi.EVAL_SKIP_CHECK).used for p in parts)
    np.testing.assert_array_equal(used, whole.used)
