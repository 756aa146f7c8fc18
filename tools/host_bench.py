"""Host-layer timings without a GPU: kt_host.cc built against the no-op engine test double (tests/host_stub/engine_stub.cc) in
its timing mode (KT_STUB_PASS_OK=1: every device pass "succeeds" with all-zero results).  What is measured is the host layer's
OWN work per call -- JSON in and out, PodRequestResourceList, dictionaries, packing, status bookkeeping -- i.e. what surrounds
the device pass in a plugin-level call.  Results are meaningless as decisions (nothing is evaluated).

    python tools/host_bench.py [pods=100000] [throttles=1000]
"""
import ctypes
import json
import os
import random
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["KT_STUB_PASS_OK"] = "1"

from kube_throttler_b200 import host  # noqa: E402
from test_scenarios import SCHED, THROTTLER, namespace, pod, throttle  # noqa: E402


def main():
    n_pods = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    n_thr = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    out = os.path.join(ROOT, "tests", "_build", "libkt_hoststub_bench.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "kube_throttler_b200", "csrc", "kt_host.cc"),
                    os.path.join(ROOT, "tests", "host_stub", "engine_stub.cc")], check=True)
    rng = random.Random(1)
    w = host.Plugin(THROTTLER, SCHED, library=ctypes.CDLL(out))
    L, h = w._L, w._h
    nss = [f"ns{i}" for i in range(50)]
    w.apply(*[namespace(n, {"team": "a"}) for n in nss])
    w.apply(*[throttle(rng.choice(nss), f"t{i}", {"app": f"a{i % 64}"}, cpu="100") for i in range(n_thr)])
    manifests = [json.dumps(pod(rng.choice(nss), f"p{i}", "100m", {"app": f"a{rng.randrange(64)}", "tier": f"t{rng.randrange(8)}"}, node="n",
                                phase="Running", requests={"memory": "128Mi"})).encode() for i in range(n_pods)]

    def timed(f, n):
        f()
        t0 = time.perf_counter()
        for _ in range(n):
            f()
        return (time.perf_counter() - t0) / n

    res = {}
    t0 = time.perf_counter()
    for m in manifests:
        L.kth_apply(h, m)
    res["kth_apply_us_per_pod_event"] = (time.perf_counter() - t0) / n_pods * 1e6
    t0 = time.perf_counter()
    L.kth_reconcile_all(h, b"2026-01-01T00:00:00Z")
    res["first_sync_ms"] = (time.perf_counter() - t0) * 1e3
    for m in manifests[:100]:
        L.kth_apply(h, m)
    t0 = time.perf_counter()
    L.kth_reconcile_all(h, b"2026-01-01T00:00:00Z")
    res["reconcile_all_after_100_events_ms"] = (time.perf_counter() - t0) * 1e3
    res["reconcile_all_ms"] = timed(lambda: L.kth_reconcile_all(h, b"2026-01-01T00:00:00Z"), 20) * 1e3
    res["metrics_scrape_ms"] = timed(lambda: L.kth_metrics(h), 5) * 1e3
    pending = [pod(rng.choice(nss), f"q{i}", "100m", {"app": f"a{rng.randrange(64)}"}, requests={"memory": "64Mi"}) for i in range(1000)]
    one = json.dumps(pending[0]).encode()
    res["pre_filter_us"] = timed(lambda: L.kth_pre_filter(h, one), 200) * 1e6
    batch = json.dumps(pending).encode()
    res["pre_filter_batch_us_per_pod"] = timed(lambda: L.kth_pre_filter_batch(h, batch), 10) * 1e6 / len(pending)
    q150 = json.dumps(pending[:150]).encode()
    res["admit_queue_150_ms"] = timed(lambda: L.kth_admit_queue(h, q150), 5) * 1e3
    res["reserve_unreserve_us"] = timed(lambda: (L.kth_reserve(h, one), L.kth_unreserve(h, one)), 100) * 1e6
    print(json.dumps({"pods": n_pods, "throttles": n_thr, **{k: round(v, 3) for k, v in res.items()}}))


if __name__ == "__main__":
    main()
