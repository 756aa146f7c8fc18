"""More seeds of the seed-parametrised tests of tests/test_plugin_vs_oracle.py on the CPU (host layer over the oracle-backed engine
test double vs the object-level oracle): test_random_world, test_resident_queue_by_key, test_admit_queue_equals_one_pod_per_cycle.
    python tools/seeds_host.py [first=100] [last=300]"""
import ctypes
import functools
import os
import subprocess
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ko  # noqa: E402
from kube_throttler_b200 import host  # noqa: E402
import test_plugin_vs_oracle as T  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    ko.build()
    out = os.path.join(ROOT, "tests", "_build", "libkt_hostoracle.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "kube_throttler_b200", "csrc", "kt_host.cc"),
                    os.path.join(ROOT, "tests", "host_stub", "engine_oracle.cc"), "-L", os.path.join(ROOT, "oracle"), "-lkt_oracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    ctor = functools.partial(host.Plugin, library=ctypes.CDLL(out))
    bad = {}
    for fn in (T.test_random_world, T.test_resident_queue_by_key, T.test_admit_queue_equals_one_pod_per_cycle):
        n_bad = 0
        for seed in range(first, last):
            try:
                fn(ko, ctor, seed)
            except IndexError:  # a scenario without enough admitted pods for the test's own script: not a parity statement
                pass
            except AssertionError:
                # the tests also assert that their scenario is not degenerate and what the verdict cache's statistics look like for
                # the seeds they were written for: only parity statements count here
                frames = [f for f in traceback.extract_tb(sys.exc_info()[2]) if "test_plugin_vs_oracle" in f.filename]
                line = frames[-1].line if frames else ""
                if any(mark in line for mark in ('["hits"]', "assert admitted", "queued >", '["rounds"]', "> 5")):
                    continue
                n_bad += 1
                print(fn.__name__, "seed", seed, "DIFFERS at:", line[:300])
        bad[fn.__name__] = n_bad
    print(f"seeds {first}..{last - 1}:", bad)
    return 1 if any(bad.values()) else 0


if __name__ == "__main__":
    sys.exit(main())
