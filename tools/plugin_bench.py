"""Plugin-level timings on a GPU: the reference-facing surface of include/kt_host.h (kth_new_plugin / kth_apply / kth_reconcile_all /
kth_pre_filter / kth_pre_filter_batch / kth_admit_queue) driven with Kubernetes manifests, as the Go plugin would drive it.

    python tools/plugin_bench.py            (prints the JSON object bench.py embeds as `e2e_plugin`)

Two worlds: BASELINE config 1 (example/throttle.yaml: 1 Throttle, 10 running pods, cpu-only threshold) for the per-call
latency of PreFilter -- beside the CPU restatement's latency for the same call -- and a C2-shaped world (50 namespaces,
1000 Throttles, 100k running pods, 10k pending pods, 4 resources) for the batched calls.
"""
import json
import os
import random
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

THROTTLER, SCHED = "kube-throttler", "my-scheduler"
NOW = "2026-01-01T00:00:00Z"


def namespace(name, labels=None):
    return {"kind": "Namespace", "metadata": {"name": name, "labels": dict(labels or {}, **{"kubernetes.io/metadata.name": name})}}


def pod(ns, name, labels, requests, node="", phase="Pending"):
    return {"kind": "Pod", "metadata": {"namespace": ns, "name": name, "labels": labels},
            "spec": {"schedulerName": SCHED, "nodeName": node, "containers": [{"name": "c", "resources": {"requests": requests}}]},
            "status": {"phase": phase}}


def throttle(ns, name, match, threshold):
    return {"kind": "Throttle", "metadata": {"namespace": ns, "name": name},
            "spec": {"throttlerName": THROTTLER, "threshold": threshold, "selector": {"selectorTerms": [{"podSelector": {"matchLabels": match}}]}}}


def timed(f, n, warm=2):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), min(ts)


def c1_latency(device):
    """BASELINE config 1: PreFilter latency per call (admit / insufficient / exceeds), GPU plugin beside the CPU restatement."""
    from kube_throttler_b200 import host
    from oracle import ko

    out = {}
    for label, w in (("b200_plugin", host.Plugin(THROTTLER, SCHED, device=device)), ("cpu_port", ko.World(THROTTLER, SCHED))):
        w.apply(namespace("default"), throttle("default", "t1", {"throttle": "t1"}, {"resourceRequests": {"cpu": "200m"}}))
        w.apply(*[pod("default", f"r{i}", {"throttle": "t1"}, {"cpu": "10m"}, node="n", phase="Running") for i in range(10)])
        w.reconcile_all(NOW)
        codes, lat = [], []
        for cpu in ("100m", "101m", "300m"):
            p = pod("default", "pending", {"throttle": "t1"}, {"cpu": cpu})
            codes.append(w.prefilter(p)["reasons"])
            med, best = timed(lambda: w.prefilter(p), 200, warm=20)
            lat.append(med * 1e6)
        out[label] = {"prefilter_us_median": [round(x, 2) for x in lat], "reasons": codes}
        if label == "b200_plugin":  # the same pod delivered by the informer first: addressed by key, verdict served from the queue pass
            p = pod("default", "pending", {"throttle": "t1"}, {"cpu": "100m"})
            w.apply(p)
            t0 = time.perf_counter()
            first = w.prefilter_key("default", "pending")
            t_first = time.perf_counter() - t0
            med, best = timed(lambda: w.prefilter_key("default", "pending"), 500, warm=20)
            out[label]["prefilter_key_us"] = {"first_call_with_its_device_pass": round(t_first * 1e6, 2), "cached_median": round(med * 1e6, 2),
                                              "reasons": first["reasons"]}
        w.close()
    assert out["b200_plugin"]["reasons"] == out["cpu_port"]["reasons"], out
    return out


def c2_world(device, n_ns=50, n_thr=1000, n_run=100_000, n_pend=10_000, seed=2):
    from kube_throttler_b200 import host

    rng = random.Random(seed)
    w = host.Plugin(THROTTLER, SCHED, device=device)
    nss = [f"ns{i}" for i in range(n_ns)]
    w.apply(*[namespace(n, {"team": f"t{i % 7}"}) for i, n in enumerate(nss)])
    for i in range(n_thr):
        thr = {"resourceRequests": {"cpu": str(rng.randrange(50, 400)), "memory": f"{rng.randrange(64, 2048)}Gi"}}
        if rng.random() < 0.5:
            thr["resourceCounts"] = {"pod": rng.randrange(20, 400)}
        w.apply(throttle(nss[i % n_ns], f"t{i}", {"app": f"a{rng.randrange(64)}"} if rng.random() < 0.8 else {"app": f"a{rng.randrange(64)}", "tier": f"x{rng.randrange(8)}"}, thr))

    def mk(i, pending):
        reqs = {"cpu": f"{rng.randrange(1, 80) * 50}m", "memory": f"{1 << rng.randrange(6, 14)}Mi"}
        if rng.random() < 0.3:
            reqs["nvidia.com/gpu"] = str(rng.choice([1, 2, 4, 8]))
        if rng.random() < 0.5:
            reqs["ephemeral-storage"] = f"{rng.randrange(1, 100)}Gi"
        labels = {"app": f"a{rng.randrange(64)}", "tier": f"x{rng.randrange(8)}", "rel": f"r{rng.randrange(16)}"}
        return pod(rng.choice(nss), f"{'q' if pending else 'p'}{i}", labels, reqs, node="" if pending else "n", phase="Pending" if pending else "Running")

    running = [json.dumps(mk(i, False)).encode() for i in range(n_run)]
    pending = [mk(i, True) for i in range(n_pend)]
    return w, running, pending


def run(device=0):
    res = {"c1": c1_latency(device)}
    w, running, pending = c2_world(device)
    L, h = w._L, w._h
    t0 = time.perf_counter()
    for m in running:
        L.kth_apply(h, m)
    res["kth_apply_us_per_pod_event"] = (time.perf_counter() - t0) / len(running) * 1e6
    t0 = time.perf_counter()
    r = w.reconcile_all(NOW)
    res["first_reconcile_all_ms"] = (time.perf_counter() - t0) * 1e3
    res["reconciled"] = r["reconciled"]
    for m in running[:100]:
        L.kth_apply(h, m)
    med, best = timed(lambda: w.reconcile_all(NOW), 5, warm=1)
    res["kth_reconcile_all_ms"] = med * 1e3
    n_thr = r["reconciled"]
    one = pending[0]
    med, best = timed(lambda: w.prefilter(one), 50, warm=5)
    res["kth_pre_filter_us"] = med * 1e6
    batch_json = json.dumps(pending).encode()
    med, best = timed(lambda: L.kth_pre_filter_batch(h, batch_json), 3, warm=1)
    res["kth_pre_filter_batch"] = {"pods": len(pending), "ms": med * 1e3, "checks_per_s": len(pending) * n_thr / med}
    # the resident queue: the informer delivers the pending pods once; PreFilter is then addressed by key / for the whole queue
    t0 = time.perf_counter()
    for m in pending:
        L.kth_apply(h, json.dumps(m).encode())
    res["kth_apply_us_per_pending_pod"] = (time.perf_counter() - t0) / len(pending) * 1e6
    import numpy as np

    verdicts = np.zeros(len(pending) + 4096, np.uint8)

    def queue_pass():
        # a status change of one throttle voids the cached verdicts of the pods it affects: the next call runs a device pass
        L.kth_apply(h, touch[queue_pass.i % len(touch)])
        queue_pass.i += 1
        return L.kth_pre_filter_queue(h, verdicts.ctypes.data, verdicts.shape[0])

    # (re-applying a Throttle manifest unchanged still recompiles the tables: the worst case for the queue pass)
    touch = [json.dumps(throttle("ns0", "t0", {"app": "a1"}, {"resourceRequests": {"cpu": str(100 + i)}})).encode() for i in range(8)]
    queue_pass.i = 0
    n_rows = L.kth_pre_filter_queue(h, verdicts.ctypes.data, verdicts.shape[0])
    med, best = timed(queue_pass, 10, warm=2)
    res["kth_pre_filter_queue"] = {"pods": int((verdicts[:n_rows] != 0).sum()), "ms_after_a_throttle_edit": med * 1e3,
                                   "checks_per_s": float((verdicts[:n_rows] != 0).sum()) * n_thr / med,
                                   "admitted": int((verdicts[:n_rows] == 1).sum())}
    med, best = timed(lambda: L.kth_pre_filter_queue(h, verdicts.ctypes.data, verdicts.shape[0]), 20, warm=2)
    res["kth_pre_filter_queue"]["ms_cached"] = med * 1e3
    # the common invalidation: a Reserve / Unreserve changed some throttles' reservations -> one pass, no table work
    rk = [(m["metadata"]["namespace"].encode(), m["metadata"]["name"].encode()) for m in pending[-64:]]

    def queue_pass_after_reserve():
        ns, name = rk[queue_pass_after_reserve.i % len(rk)]
        (L.kth_reserve_key if (queue_pass_after_reserve.i // len(rk)) % 2 == 0 else L.kth_unreserve_key)(h, ns, name)
        queue_pass_after_reserve.i += 1
        return L.kth_pre_filter_queue(h, verdicts.ctypes.data, verdicts.shape[0])

    queue_pass_after_reserve.i = 0
    med, best = timed(queue_pass_after_reserve, 20, warm=3)
    n_q = float((verdicts[:n_rows] != 0).sum())
    res["kth_pre_filter_queue"]["ms_after_a_reserve"] = med * 1e3
    res["kth_pre_filter_queue"]["checks_per_s_after_a_reserve"] = n_q * n_thr / med
    keys = [(m["metadata"]["namespace"].encode(), m["metadata"]["name"].encode()) for m in pending[:2000]]
    t0 = time.perf_counter()
    for ns, name in keys:
        L.kth_pre_filter_key(h, ns, name)
    res["kth_pre_filter_key_us_cached"] = (time.perf_counter() - t0) / len(keys) * 1e6
    # the scheduler's cycle by key: PreFilter -> Reserve on Success; a Reserve voids the verdicts of pods sharing a throttle
    stats0 = w.queue_stats()
    t0 = time.perf_counter()
    admitted = 0
    for ns, name in keys[:500]:
        r = L.kth_pre_filter_key(h, ns, name)
        if r.startswith(b'{"code":"Success"'):
            L.kth_reserve_key(h, ns, name)
            admitted += 1
    stats1 = w.queue_stats()
    res["scheduling_cycle_by_key"] = {"pods": 500, "us_per_pod": (time.perf_counter() - t0) / 500 * 1e6, "admitted": admitted,
                                      "device_passes": stats1["passes"] - stats0["passes"], "cache_hits": stats1["hits"] - stats0["hits"]}
    # (pods that hold no reservation yet: a queue with pods reserved in the cycle above takes the host's pod-by-pod passes --
    # Reserve is idempotent per pod, the device's prefix sums would count them twice)
    queue = pending[2000:3000]
    t0 = time.perf_counter()
    adm = w.admit_queue(queue)
    res["kth_admit_queue"] = {"pods": len(queue), "ms": (time.perf_counter() - t0) * 1e3, "rounds": adm["rounds"], "admitted": adm["admitted"],
                              "on_device": bool(adm.get("onDevice", False))}
    w.close()
    return res


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 0)))
