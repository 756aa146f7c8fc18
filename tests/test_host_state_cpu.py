"""The host layer's bookkeeping (kt_host.cc: informer events, limits, dictionaries, status JSON, gauges) on the CPU.

kt_host.cc is compiled here against a TEST DOUBLE of the device engine (tests/host_stub/engine_stub.cc) that accepts uploads
and fails every call that would need a device pass -- nothing is evaluated on the CPU, and a test that strays into PreFilter /
reconcile gets the engine's error.  What this covers without a GPU: which manifests are accepted or refused, that a refused
manifest leaves no trace, status round trips through Apply, tombstones, gauges of applied statuses."""
import ctypes as C
import json
import os
import subprocess

import pytest

from test_scenarios import SCHED, THROTTLER, namespace, pod, throttle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRCS = [os.path.join(ROOT, "kube_throttler_b200", "csrc", "kt_host.cc"), os.path.join(ROOT, "tests", "host_stub", "engine_stub.cc")]
DEPS = SRCS + [os.path.join(ROOT, "kube_throttler_b200", "csrc", f) for f in ("kt_json.h", "kt_quantity.h")] + \
    [os.path.join(ROOT, "include", f) for f in ("kt_b200.h", "kt_host.h")]
OUT = os.path.join(ROOT, "tests", "_build", "libkt_hoststub.so")


@pytest.fixture(scope="module")
def stub():
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(f) for f in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-o", OUT] + SRCS, check=True)
    L = C.CDLL(OUT)
    vp, cp = C.c_void_p, C.c_char_p
    L.kth_new_plugin.argtypes = [C.POINTER(vp), cp, C.c_int]
    L.kth_free.argtypes = [vp]
    L.kth_free.restype = None
    for name, args in (("kth_apply", [vp, cp]), ("kth_delete", [vp, cp, cp, cp]), ("kth_get_status", [vp, cp, cp]), ("kth_metrics", [vp]),
                       ("kth_pre_filter", [vp, cp]), ("kth_reconcile_all", [vp, cp]), ("kth_reserved", [vp, C.c_int, cp])):
        getattr(L, name).argtypes = args
        getattr(L, name).restype = cp
    return L


class World:
    def __init__(self, L):
        self.L, self.h = L, C.c_void_p()
        assert L.kth_new_plugin(C.byref(self.h), json.dumps({"name": THROTTLER, "targetSchedulerName": SCHED}).encode(), 0) == 0

    def _res(self, raw):
        out = json.loads(raw.decode())
        if isinstance(out, dict) and set(out) == {"error"}:
            raise RuntimeError(out["error"])
        return out

    def apply(self, *manifests):
        for m in manifests:
            self._res(self.L.kth_apply(self.h, json.dumps(m).encode()))

    def delete(self, kind, name, ns=""):
        return self._res(self.L.kth_delete(self.h, kind.encode(), ns.encode(), name.encode()))

    def status(self, name, ns=""):
        return self._res(self.L.kth_get_status(self.h, ns.encode(), name.encode()))

    def prefilter(self, p):
        return self._res(self.L.kth_pre_filter(self.h, json.dumps(p).encode()))

    def close(self):
        self.L.kth_free(self.h)


@pytest.fixture
def world(stub):
    w = World(stub)
    yield w
    w.close()


STATUS = {"calculatedThreshold": {"threshold": {"resourceRequests": {"cpu": "700m"}}, "calculatedAt": "2026-01-01T00:00:00Z", "messages": ["m"]},
          "throttled": {"resourceCounts": {"pod": False}, "resourceRequests": {"cpu": True}},
          "used": {"resourceCounts": {"pod": 3}, "resourceRequests": {"cpu": "1500m", "memory": "512Mi"}}}


def test_status_round_trips_and_survives_a_spec_update(world):
    t = throttle("default", "t", {"a": "1"}, pod_cnt=5, cpu="1")
    world.apply(namespace("default"), dict(t, status=STATUS))
    s = world.status("t", "default")
    assert s["calculatedThreshold"]["threshold"] == {"resourceRequests": {"cpu": "0.7"}} and s["calculatedThreshold"]["calculatedAtSet"] is True
    assert s["calculatedThreshold"]["messages"] == ["m"] and s["throttled"] == {"resourceCounts": {"pod": False}, "resourceRequests": {"cpu": True}}
    assert s["used"] == {"resourceCounts": {"pod": 3}, "resourceRequests": {"cpu": "1.5", "memory": "536870912"}}
    world.apply(throttle("default", "t", {"a": "2"}, cpu="2"))        # spec update without a status: the status subresource is kept
    assert world.status("t", "default") == s
    world.apply(dict(throttle("default", "t", {"a": "2"}, cpu="2"), status={}))  # an explicit (empty) status replaces it
    s2 = world.status("t", "default")
    assert s2["used"] == {} and s2["calculatedThreshold"]["calculatedAtSet"] is False and "messages" not in s2["calculatedThreshold"]


def test_refused_manifests_leave_no_trace(world):
    world.apply(namespace("default"), throttle("default", "t", {"a": "1"}, cpu="1"))
    before = world.status("t", "default")
    bad_time = dict(throttle("default", "t", {"a": "1"}, cpu="3", extra={"example.com/only-here": "1"}),
                    status=dict(STATUS, calculatedThreshold={"threshold": {}, "calculatedAt": "yesterday"}))
    with pytest.raises(RuntimeError, match="parsing time"):
        world.apply(bad_time)
    assert world.status("t", "default") == before                     # neither the new spec nor its status was committed
    with pytest.raises(RuntimeError, match="quantit"):
        world.apply(throttle("default", "t2", {"a": "1"}, cpu="1x"))
    with pytest.raises(RuntimeError, match="not found"):
        world.status("t2", "default")
    with pytest.raises(RuntimeError, match="more than 32 labels"):
        world.apply(pod("default", "fat", "100m", {f"k{i}": "v" for i in range(33)}))
    with pytest.raises(RuntimeError, match="more than 31 labels"):
        world.apply(namespace("fat-ns", {f"l{i}": "x" for i in range(33)}))
    # a 32nd resource name is refused -- and the 30 names the refused pod had already brought along go with it: afterwards
    # thirty OTHER new names still fit
    with pytest.raises(RuntimeError, match="distinct resource names"):
        world.apply(pod("default", "greedy", "100m", {"a": "1"}, requests={f"example.com/r{i}": "1" for i in range(40)}))
    world.apply(pod("default", "modest", "100m", {"a": "1"}, requests={f"example.com/s{i}": "1" for i in range(30)}))
    with pytest.raises(RuntimeError, match="distinct resource names"):
        world.apply(pod("default", "one-more", "100m", {"a": "1"}, requests={"example.com/the-32nd": "1"}))
    with pytest.raises(RuntimeError, match="unsupported kind"):
        world.apply({"kind": "Deployment", "metadata": {"name": "d"}})


def test_device_dependent_calls_fail_loudly_on_the_stub(world):
    """Nothing under kt_host.cc can answer a PreFilter or run a reconcile without the engine."""
    world.apply(namespace("default"), throttle("default", "t", {"a": "1"}, cpu="1"), pod("default", "p0", "100m", {"a": "1"}, node="n", phase="Running"))
    with pytest.raises(RuntimeError, match="kt_evaluate"):
        world.prefilter(pod("default", "x", "100m", {"a": "1"}))
    with pytest.raises(RuntimeError, match="kt_evaluate"):
        world._res(world.L.kth_reconcile_all(world.h, b"2026-01-01T00:00:00Z"))


def test_gauges_are_only_recorded_by_reconcile(world):
    world.apply(namespace("default"), dict(throttle("default", "t", {"a": "1"}, cpu="1"), status=STATUS))
    assert world.L.kth_metrics(world.h).decode() == ""  # an applied status is the informer's copy, not a reconcile


def test_pod_delete_and_reapply_reuse_rows(world):
    world.apply(namespace("default"))
    for i in range(4):
        world.apply(pod("default", f"p{i}", "100m", {"a": "1"}, node="n", phase="Running"))
    assert world.delete("Pod", "p1", "default") == {"ok": True}
    assert world.delete("Pod", "nope", "default") == {"ok": True}   # deleting what is not there is not an error (DeleteFunc of a stale key)
    world.apply(pod("default", "p9", "1", {"a": "1"}, node="n", phase="Running"), pod("default", "p0", "200m", {"a": "2"}, node="n", phase="Running"))
    assert world.delete("Throttle", "nope", "default") == {"ok": True}


def test_column_overflow_is_proved_incrementally(host_on_oracle):
    """The packer refuses a snapshot whose column sum could wrap int64 (sum of |v| >= 2^62).  The per-column totals follow the
    pod events, so the proof does not walk the pod table at every sync; it has to notice an overflow that an update or a new
    pod brings about and to forget it when the pod goes away.  (Host layer over the oracle-backed engine double.)"""
    w = host_on_oracle(THROTTLER, SCHED)
    w.apply(namespace("default"), throttle("default", "t", {"a": "1"}, cpu="1"))
    w.apply(pod("default", "p0", "100m", {"a": "1"}, node="n", phase="Running", requests={"memory": "2Ei"}))
    assert w.reconcile_all("2026-01-01T00:00:00Z")["reconciled"] == 1
    w.apply(pod("default", "p1", "100m", {"a": "1"}, node="n", phase="Running", requests={"memory": "2Ei"}))  # 2^61 + 2^61
    with pytest.raises(RuntimeError, match="can overflow int64"):
        w.reconcile_all("2026-01-01T00:00:00Z")
    w.apply(pod("default", "p1", "100m", {"a": "1"}, node="n", phase="Running", requests={"memory": "1Ei"}))  # update shrinks it
    assert w.reconcile_all("2026-01-01T00:00:00Z")["reconciled"] == 1
    w.apply(pod("default", "p2", "100m", {"a": "1"}, node="n", phase="Running", requests={"memory": "1Ei"}))  # 2^61 + 2^60 + 2^60
    with pytest.raises(RuntimeError, match="can overflow int64"):
        w.reconcile_all("2026-01-01T00:00:00Z")
    w.delete("Pod", "p0", "default")
    assert w.reconcile_all("2026-01-01T00:00:00Z")["reconciled"] == 1
    assert w.status("t", "default")["used"]["resourceRequests"]["memory"] == str(2**61)
    w.close()


def test_status_manifest_is_what_update_status_would_send(host_on_oracle):
    """kth_get_status_manifest: encoding/json of v1alpha1.ThrottleStatus -- declaration order, sorted map keys, canonical
    quantities (the integration suite asserts "500m", "1", ... through Quantity.String, util_throttle_test.go:169-177), RFC3339."""
    w = host_on_oracle(THROTTLER, SCHED)
    w.apply(namespace("default"), throttle("default", "t", {"a": "1"}, pod_cnt=2, cpu="1", extra={"memory": "1Gi"}),
            throttle("default", "empty", {"a": "zzz"}, cpu="200m"))
    assert w.status_manifest("t", "default") == '{"calculatedThreshold":{"threshold":{},"calculatedAt":null},"throttled":{"resourceCounts":{"pod":false}},"used":{}}'
    w.apply(pod("default", "p0", "250m", {"a": "1"}, node="n", phase="Running", requests={"memory": "512Mi"}),
            pod("default", "p1", "250m", {"a": "1"}, node="n", phase="Running", requests={"memory": "512Mi", "example.com/x": "1500m"}))
    w.reconcile_all("2026-01-01T00:00:00Z")
    assert w.status_manifest("t", "default") == (
        '{"calculatedThreshold":{"threshold":{"resourceCounts":{"pod":2},"resourceRequests":{"cpu":"1","memory":"1Gi"}},"calculatedAt":"2026-01-01T00:00:00Z"},'
        '"throttled":{"resourceCounts":{"pod":true},"resourceRequests":{"cpu":false,"memory":true}},'
        '"used":{"resourceCounts":{"pod":2},"resourceRequests":{"cpu":"500m","example.com/x":"1500m","memory":"1Gi"}}}')
    # no matched pod: used stays the zero ResourceAmount (Q3), the threshold map has one entry
    assert w.status_manifest("empty", "default") == (
        '{"calculatedThreshold":{"threshold":{"resourceRequests":{"cpu":"200m"}},"calculatedAt":"2026-01-01T00:00:00Z"},'
        '"throttled":{"resourceCounts":{"pod":false},"resourceRequests":{"cpu":false}},"used":{}}')
    # the suite's spellings: 20 x 50m == "1", 900m stays "900m"
    w.apply(throttle("default", "u", {"b": "1"}, cpu="1"))
    for i in range(20):
        w.apply(pod("default", f"q{i}", "50m", {"b": "1"}, node="n", phase="Running"))
    w.reconcile_all("2019-02-01T09:00:00+09:00")
    m = json.loads(w.status_manifest("u", "default"))
    assert m["used"] == {"resourceCounts": {"pod": 20}, "resourceRequests": {"cpu": "1"}} and m["calculatedThreshold"]["calculatedAt"] == "2019-02-01T00:00:00Z"
    w.close()


def test_pod_update_is_committed_even_when_its_follow_up_pass_is_refused(host_on_oracle):
    """An informer update cannot be refused half-way: the pod is stored BEFORE the pass that moves its reservations, so when that
    pass throws (here: the snapshot's memory column can overflow) the apply succeeds with a warning and the resource names the
    new object brought stay interned -- rolling them back would leave the stored pod pointing at column ids that a later
    resource name re-uses (`foo` aliased to `bar`)."""
    w = host_on_oracle(THROTTLER, SCHED)
    w.apply(namespace("default"), throttle("default", "t", {"a": "1"}, cpu="1"))
    w.apply(pod("default", "p0", "100m", {"a": "1"}, node="n", phase="Running", requests={"memory": "2Ei"}))
    w.apply(pod("default", "p1", "100m", {"a": "1"}, node="n", phase="Running", requests={"memory": "2Ei"}))
    out = w.apply(pod("default", "p1", "100m", {"a": "2"}, node="n", phase="Running", requests={"memory": "2Ei", "example.com/foo": "1"}))
    assert out["ok"] is True and "can overflow int64" in out["warning"]
    w.delete("Pod", "p0", "default")  # the snapshot is provable again
    bar = throttle("default", "tbar", {"a": "2"}, cpu="1")
    bar["spec"]["threshold"]["resourceRequests"] = {"example.com/bar": "5"}
    w.apply(bar)
    assert w.reconcile_all("2026-01-01T00:00:00Z")["reconciled"] == 2
    used = w.status("tbar", "default")["used"]
    assert used["resourceCounts"]["pod"] == 1
    assert used["resourceRequests"]["example.com/foo"] == "1" and "example.com/bar" not in used["resourceRequests"]
    w.close()


def test_deleted_pod_never_keeps_a_reservation(host_on_oracle):
    """DeleteFunc un-reserves the pod from its affected throttles (throttle_controller.go:509-515).  When the pass that would name
    them cannot run, the pod still must not stay in the reservation cache (reconcile only un-reserves pods it can still see)."""
    w = host_on_oracle(THROTTLER, SCHED)
    w.apply(namespace("default"), throttle("default", "t", {"a": "1"}, cpu="10"))
    w.reconcile_all("2026-01-01T00:00:00Z")
    q0 = pod("default", "q0", "1", {"a": "1"})
    assert w.prefilter(q0)["code"] == "Success" and w.reserve(q0)["code"] == "Success"
    w.apply(dict(q0, spec=dict(q0["spec"], nodeName="n")))  # scheduled, not yet observed by a reconcile: still reserved
    assert "default/q0" in w.reserved("Throttle", "default/t")["pods"]
    w.apply(pod("default", "p0", "100m", {"b": "1"}, node="n", phase="Running", requests={"memory": "2Ei"}),
            pod("default", "p1", "100m", {"b": "1"}, node="n", phase="Running", requests={"memory": "2Ei"}))
    with pytest.raises(RuntimeError, match="can overflow int64"):
        w.delete("Pod", "q0", "default")
    w.delete("Pod", "p1", "default")
    assert "default/q0" not in w.reserved("Throttle", "default/t")["pods"]
    assert w.prefilter(pod("default", "q1", "9500m", {"a": "1"}))["code"] == "Success"  # nothing reserved any more
    w.close()


def test_rows_stay_namespace_clustered_under_churn(host_on_oracle):
    """The device kernels walk a warp's 32 rows word by word, one or two rounds when they share a namespace: the host hands rows
    out from per-namespace arenas of 32 consecutive rows, and a deleted pod's row goes back to its namespace, so after any mix of
    arrivals and departures every 32-row chunk still holds pods of ONE namespace (the reference's pod informer is indexed by
    namespace, plugin.go:81-84)."""
    import random

    rng = random.Random(4)
    w = host_on_oracle(THROTTLER, SCHED)
    nss = [f"ns{i}" for i in range(7)]
    w.apply(*[namespace(n) for n in nss])
    live = {}
    for step in range(3000):
        if live and rng.random() < 0.4:
            key = rng.choice(sorted(live))
            w.delete("Pod", key[1], key[0])
            del live[key]
        else:
            key = (rng.choice(nss), f"p{step}")
            w.apply(pod(key[0], key[1], "100m", {"a": "1"}, node="n", phase="Running"))
            live[key] = True
    chunk_ns = {}
    for ns_, name in live:
        row = w.pod_row(ns_, name)
        assert row >= 0
        assert chunk_ns.setdefault(row // 32, ns_) == ns_, (row, ns_, chunk_ns[row // 32])
    rows = sorted(w.pod_row(*k) for k in live)
    assert len(set(rows)) == len(rows) and rows[-1] < len(live) + 32 * len(nss) + 32 * 40  # slots are reused, the table stays compact
    w.close()


def test_deleted_throttles_give_their_device_columns_back(host_on_oracle):
    """A scheduler that lives for months sees throttles come and go: the device column of a deleted throttle is handed to the
    next new one (M, the table compile and every pass would only ever grow otherwise) -- and the lists the caller sees stay in
    CREATION order, as the reference's lister-backed slices are, not in column order: the names inside a PreFilter reason."""
    w = host_on_oracle(THROTTLER, SCHED)
    w.apply(namespace("default"))
    w.apply(*[throttle("default", f"t{i}", {"a": "1"}, cpu="100m") for i in range(5)])
    w.reconcile_all()
    assert w.queue_stats()["throttleColumns"] == 5
    for round_ in range(20):  # churn: three go, three come, twenty times over
        for i in (1, 2, 3):
            w.delete("Throttle", f"t{i}" if round_ == 0 else f"n{round_ - 1}_{i}", "default")
        w.apply(*[throttle("default", f"n{round_}_{i}", {"a": "1"}, cpu="100m") for i in (1, 2, 3)])
        w.reconcile_all()
    st = w.queue_stats()
    assert st["throttleColumns"] == 5 and st["liveThrottles"] == 5, st
    # t0 and t4 are the oldest, then the last round's three in the order they were applied -- whatever columns they sit in
    r = w.prefilter(pod("default", "x", "500m", {"a": "1"}))
    assert r["reasons"] == ["throttle[pod-requests-exceeds-threshold]=default/t0,default/t4,default/n19_1,default/n19_2,default/n19_3"], r
    w.delete("Throttle", "t0", "default")
    w.apply(throttle("default", "t0", {"a": "1"}, cpu="100m"))  # back under its old name, in its old column: now the youngest
    r = w.prefilter(pod("default", "x", "500m", {"a": "1"}))
    assert r["reasons"] == ["throttle[pod-requests-exceeds-threshold]=default/t4,default/n19_1,default/n19_2,default/n19_3,default/t0"], r
    assert w.queue_stats()["throttleColumns"] == 5
    w.close()


def test_label_dictionaries_hold_what_selectors_mention(oracle, host_on_oracle):
    """Pods come and go with an unbounded stream of label values (pod-template-hash, controller-uid, job-name ...): the label
    dictionaries -- and the device's value tables -- hold what the SELECTORS mention, not every value ever seen.  A label no selector
    can see is dropped, a mentioned key with an unmentioned value is "some other value"; when a later throttle mentions a value
    that pods already carry, exactly those rows are packed again and match."""
    ref, dut = oracle.World(THROTTLER, SCHED), host_on_oracle(THROTTLER, SCHED)
    both = lambda *m: (ref.apply(*m), dut.apply(*m))
    both(namespace("default"))
    both(throttle("default", "t-app", {"app": "web"}, pod_cnt=100))
    both({"kind": "Throttle", "metadata": {"namespace": "default", "name": "t-notin"},
          "spec": {"throttlerName": THROTTLER, "threshold": {"resourceCounts": {"pod": 100}},
                   "selector": {"selectorTerms": [{"podSelector": {"matchExpressions": [{"key": "tier", "operator": "NotIn", "values": ["db"]},
                                                                                          {"key": "job", "operator": "Exists"}]}}]}}})
    for i in range(400):  # every pod brings values nobody has seen before
        both(pod("default", f"p{i}", "100m", {"app": "web" if i % 3 == 0 else f"app-{i}", "pod-template-hash": f"h{i}", "job": f"job-{i}", "tier": f"tier-{i % 7}"},
                 node="n", phase="Running"))
        if i % 2:
            ref.delete("Pod", f"p{i - 1}", "default"), dut.delete("Pod", f"p{i - 1}", "default")
    ref.reconcile_all(), dut.reconcile_all()
    st = dut.queue_stats()
    assert st["labelKeys"] == 3 and st["labelValues"] == 3 + 2, st  # app / tier / job; "web", "db" + one "other value" per key
    for name in ("t-app", "t-notin"):
        assert dut.status(name, "default")["used"] == ref.status(name, "default")["used"], name
    assert dut.status("t-app", "default")["used"]["resourceCounts"]["pod"] > 0 and dut.status("t-notin", "default")["used"]["resourceCounts"]["pod"] == 200
    # a throttle that mentions a value (and a key) pods already carry under "other" / "invisible"
    both(throttle("default", "t-late", {"app": "app-7", "pod-template-hash": "h7"}, pod_cnt=100))
    ref.reconcile_all(), dut.reconcile_all()
    assert dut.status("t-late", "default")["used"] == ref.status("t-late", "default")["used"]
    assert dut.status("t-late", "default")["used"]["resourceCounts"] == {"pod": 1}  # p7, found under its re-packed labels
    probe = pod("default", "x", "100m", {"app": "app-7", "pod-template-hash": "h7", "tier": "db"})
    a, b = ref.prefilter(probe), dut.prefilter(probe)
    assert (a["code"], a["reasons"]) == (b["code"], b["reasons"])
    st = dut.queue_stats()
    assert st["labelKeys"] == 4 and st["labelValues"] == 4 + 4, st
    dut.close()


def test_prefilter_of_a_manifest_does_not_use_up_resource_columns(oracle, host_on_oracle):
    """PreFilter keeps no state.  A pending pod that asks for a resource nobody has a column for -- no throttle mentions it, no
    pod of the informer requests it -- cannot be influenced by it (IsThrottledFor only looks at threshold resources): the name gets
    no column, and a stream of such pods neither hits the limit of distinct resource names nor re-creates the engine.  Reserve
    does intern (the reservation carries every name)."""
    ref, dut = oracle.World(THROTTLER, SCHED), host_on_oracle(THROTTLER, SCHED)
    both = lambda *m: (ref.apply(*m), dut.apply(*m))
    both(namespace("default"))
    both(throttle("default", "t", {"a": "1"}, cpu="300m", extra={"example.com/gpu": "2"}))
    both(pod("default", "r0", "100m", {"a": "1"}, node="n", phase="Running", requests={"memory": "64Mi"}))
    ref.reconcile_all(), dut.reconcile_all()
    cols0 = dut.queue_stats()["resourceColumns"]
    assert cols0 == 3  # cpu, example.com/gpu, memory
    for i in range(60):  # far more distinct names than the 31 columns there are
        p = pod("default", f"q{i}", "150m" if i % 2 else "250m", {"a": "1"}, requests={f"vendor-{i}.example.com/widget": str(i + 1), "example.com/gpu": str(i % 4)})
        a, b = ref.prefilter(p), dut.prefilter(p)
        assert (a["code"], a["reasons"]) == (b["code"], b["reasons"]), (i, a, b)
    batch = [pod("default", f"b{i}", "100m", {"a": "1"}, requests={f"batch-{i}.example.com/x": "1"}) for i in range(40)]
    assert [(r["code"], r["reasons"]) for r in dut.prefilter_batch(batch)] == [(r["code"], r["reasons"]) for r in (ref.prefilter(p) for p in batch)]
    assert dut.queue_stats()["resourceColumns"] == cols0
    with pytest.raises(RuntimeError):  # the quantity of an unknown name is still parsed
        dut.prefilter(pod("default", "bad", "100m", {"a": "1"}, requests={"unknown.example.com/x": "1x"}))
    p = pod("default", "keep", "100m", {"a": "1"}, requests={"kept.example.com/x": "3"})
    assert ref.reserve(p)["code"] == dut.reserve(p)["code"] == "Success"
    assert dut.queue_stats()["resourceColumns"] == cols0 + 1
    a, b = ref.reserved("Throttle", "default/t"), dut.reserved("Throttle", "default/t")
    assert sorted(a["pods"]) == sorted(b["pods"]) == ["default/keep"]
    dut.close()


def test_device_columns_are_laid_out_by_namespace(oracle, host_on_oracle):
    """What a pass costs per pod is the number of 32-throttle words in which some throttle can apply to its namespace.  Throttles
    arrive in whatever order their owners create them; the host lays the device columns out by namespace (ClusterThrottles by the
    set of namespaces their namespaceSelectors admit), so a namespace's throttles share words -- and nothing the caller sees
    depends on it: the names inside a reason stay in creation order."""
    import random

    rng = random.Random(3)
    ref, dut = oracle.World(THROTTLER, SCHED), host_on_oracle(THROTTLER, SCHED)
    both = lambda *m: (ref.apply(*m), dut.apply(*m))
    nss = [f"ns{i:02d}" for i in range(24)]
    both(*[namespace(n, {"team": f"team{i % 3}"}) for i, n in enumerate(nss)])
    objs = [throttle(n, f"t{j}", {"a": "1"}, cpu="100m") for n in nss for j in range(10)]
    objs += [{"kind": "ClusterThrottle", "metadata": {"name": f"ct-{team}-{j}"},
              "spec": {"throttlerName": THROTTLER, "threshold": {"resourceRequests": {"cpu": "100m"}},
                       "selector": {"selectorTerms": [{"podSelector": {"matchLabels": {"a": "1"}}, "namespaceSelector": {"matchLabels": {"team": team}}}]}}}
             for team in ("team0", "team1", "team2") for j in range(12)]
    rng.shuffle(objs)  # 276 throttles = 9 words, created in random order: every namespace would see ~9 of them
    both(*objs)
    ref.reconcile_all(), dut.reconcile_all()
    st = dut.queue_stats()
    assert st["liveThrottles"] == st["throttleColumns"] == len(objs)
    assert st["wordsPerNamespaceX100"] <= 400, st  # 10 Throttles of its own (1-2 words) + its team's 12 ClusterThrottles (1-2 words)
    order = [o["metadata"].get("namespace", "") + "/" + o["metadata"]["name"] for o in objs]
    for n in rng.sample(nss, 6):
        probe = pod(n, "x", "500m", {"a": "1"})
        a, b = ref.prefilter(probe), dut.prefilter(probe)
        assert (a["code"], a["reasons"]) == (b["code"], b["reasons"])
        names = b["reasons"][0].split("=")[1].split(",")  # clusterthrottle[...] comes first (plugin.go:182-213)
        assert names == [x for x in order if x in names] and len(names) == 12, names
    dut.close()
