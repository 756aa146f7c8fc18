"""More seeds of tests/test_plugin_vs_oracle.py::test_event_stream_chaos and ::test_resident_queue_event_stream on the CPU: the product's host layer (kt_host.cc) over the
oracle-backed engine test double against the object-level oracle.      python tools/chaos_host.py [first=0] [last=200]"""
import ctypes
import functools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ko  # noqa: E402
from kube_throttler_b200 import host  # noqa: E402
import test_plugin_vs_oracle as T  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    ko.build()
    out = os.path.join(ROOT, "tests", "_build", "libkt_hostoracle.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "kube_throttler_b200", "csrc", "kt_host.cc"),
                    os.path.join(ROOT, "tests", "host_stub", "engine_oracle.cc"), "-L", os.path.join(ROOT, "oracle"), "-lkt_oracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    ctor = functools.partial(host.Plugin, library=ctypes.CDLL(out))
    bad = 0
    for seed in range(first, last):
        import functools as ft
        for name, stream in (("events", ft.partial(T.run_event_stream, check_every_reconcile=True)), ("queue", T.run_queue_stream), ("statuses", T.run_status_stream), ("growth", T.run_growth_stream),
                             ("events, 80 throttles x 9 namespaces", ft.partial(T.run_event_stream, n_thr=80, n_ns=9, check_every_reconcile=True)),
                             ("queue, 70 throttles, namespace deletes", ft.partial(T.run_queue_stream, n_thr=70, ns_deletes=True))):
            try:
                stream(ko, ctor, seed)
            except AssertionError as e:
                bad += 1
                print(name, "seed", seed, "DIFFERS:", str(e)[:800])
    print(f"seeds {first}..{last - 1}: {bad} differences")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
