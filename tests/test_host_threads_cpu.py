"""The plugin handle is called from several OS threads at once in a scheduler (PreFilter / Reserve on the scheduling goroutine,
Unreserve on binding goroutines, informer handlers and reconcile workers on theirs; cgo calls hop threads).  kt_host.h promises
one mutex per handle and thread-local result strings; this hammers one handle from eight threads (ctypes drops the GIL around
every call) on the oracle-backed engine double and checks that nothing is lost or garbled."""
import json
import threading

from test_scenarios import NOW, SCHED, THROTTLER, namespace, pod, throttle


def test_one_handle_many_threads(host_on_oracle):
    w = host_on_oracle(THROTTLER, SCHED)
    w.apply(namespace("default"), *[throttle("default", f"t{i}", {"app": f"a{i}"}, pod_cnt=1000, cpu="1000") for i in range(8)])
    errors, results = [], {}

    def informer(k):
        try:
            for i in range(60):
                w.apply(pod("default", f"p{k}-{i}", "100m", {"app": f"a{k}"}, node="n", phase="Running"))
                if i % 7 == 3:
                    w.delete("Pod", f"p{k}-{i - 1}", "default")
        except Exception as e:  # noqa: BLE001
            errors.append(("informer", k, repr(e)))

    def scheduler(k):
        try:
            for i in range(25):
                p = pod("default", f"q{k}-{i}", "50m", {"app": f"a{k}"})
                r = w.prefilter(p)
                assert r["code"] == "Success" and r["throttle"]["affected"] == [f"default/t{k}"], r
                assert w.reserve(p)["code"] == "Success"
                assert w.unreserve(p)["code"] == "Success"
        except Exception as e:  # noqa: BLE001
            errors.append(("scheduler", k, repr(e)))

    def controller(k):
        try:
            for _ in range(12):
                assert w.reconcile_all(NOW)["reconciled"] == 8
                s = w.status(f"t{k}", "default")
                json.loads(w.status_manifest(f"t{k}", "default"))
                results[k] = s
        except Exception as e:  # noqa: BLE001
            errors.append(("controller", k, repr(e)))

    threads = [threading.Thread(target=f, args=(k,)) for k in range(3) for f in (informer, scheduler)] + [threading.Thread(target=controller, args=(k,)) for k in (0, 5)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    w.reconcile_all(NOW)
    for k in range(3):  # 60 applied, 8 deleted (i = 3, 10, ..., 59 -> 9 deletes of i-1; the last index 59 % 7 == 3)
        deleted = len([i for i in range(60) if i % 7 == 3])
        s = w.status(f"t{k}", "default")
        assert s["used"]["resourceCounts"]["pod"] == 60 - deleted, (k, s)
    for k in range(3):
        assert w.reserved("Throttle", f"default/t{k}")["pods"] == []
    w.close()
