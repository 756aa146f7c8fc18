"""Mutation fuzzing of the host C ABI (include/kt_host.h) under ASan/UBSan, without a GPU.

kt_host.cc is built against the no-op engine test double (tests/host_stub/engine_stub.cc); valid manifests are mutated (cuts,
inserted JSON tokens, flipped bytes) and pushed through kth_apply / kth_pre_filter / kth_reserve / kth_pre_filter_batch /
kth_admit_queue / kth_eval.  Every call must come back with a string (a result or {"error": ...}); the sanitizers watch the rest.

    g++ -std=c++17 -O1 -g -fsanitize=address -fsanitize=undefined -fno-omit-frame-pointer -shared -fPIC -o /tmp/stub_asan.so \\
        kube_throttler_b200/csrc/kt_host.cc tests/host_stub/engine_stub.cc
    LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \\
        python tools/fuzz_host.py /tmp/stub_asan.so 6000
"""
import sys, json, ctypes as C, random
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_scenarios import pod, throttle, clthrottle, namespace, THROTTLER, SCHED
L = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else '/tmp/stub_asan.so')
vp, cp = C.c_void_p, C.c_char_p
L.kth_new_plugin.argtypes=[C.POINTER(vp), cp, C.c_int]
for n,a in (("kth_apply",[vp,cp]),("kth_pre_filter",[vp,cp]),("kth_pre_filter_batch",[vp,cp]),("kth_admit_queue",[vp,cp]),("kth_reserve",[vp,cp]),("kth_eval",[cp]),("kth_delete",[vp,cp,cp,cp]),("kth_get_status_manifest",[vp,cp,cp]),("kth_metrics",[vp])):
    getattr(L,n).argtypes=a; getattr(L,n).restype=cp
h=vp(); assert L.kth_new_plugin(C.byref(h), json.dumps({"name":THROTTLER,"targetSchedulerName":SCHED}).encode(), 0)==0
rng=random.Random(5)
seeds=[namespace("default",{"a":"b"}), throttle("default","t",{"a":"1"},pod_cnt=3,cpu="1",extra={"memory":"1Gi"},overrides=[{"begin":"2026-01-01T00:00:00Z","end":"","threshold":{"resourceRequests":{"cpu":"2"}}}]),
       clthrottle("c",{"x":"y"},{"a":"1"},cpu="500m"), pod("default","p","100m",{"a":"1"},node="n",phase="Running",requests={"memory":"64Mi"}),
       dict(throttle("default","s",{"a":"1"},cpu="1"), status={"calculatedThreshold":{"threshold":{"resourceRequests":{"cpu":"1"}},"calculatedAt":"2026-01-01T00:00:00Z"},"throttled":{"resourceCounts":{"pod":True},"resourceRequests":{"cpu":False}},"used":{"resourceCounts":{"pod":1},"resourceRequests":{"cpu":"100m"}}})]
for s in seeds: L.kth_apply(h, json.dumps(s).encode())
tokens=[b'"',b'{',b'}',b'[',b']',b':',b',',b'null',b'true',b'1e999',b'-',b'\\u0000',b'\\',b'9'*40,b'"cpu"',b'"1Ei"',b'"In"',b'"matchExpressions"',b'""']
n=0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6000):
    base=bytearray(json.dumps(rng.choice(seeds)).encode())
    for _ in range(rng.randrange(1,4)):
        k=rng.random(); pos=rng.randrange(len(base)+1)
        if k<0.3 and base: del base[pos:pos+rng.randrange(1,8)]
        elif k<0.7: base[pos:pos]=rng.choice(tokens)
        elif base: base[min(pos,len(base)-1)]=rng.randrange(32,127)
    data=bytes(base).replace(b"\x00",b"")
    for fn in (L.kth_apply, L.kth_pre_filter, L.kth_reserve):
        out=fn(h,data); n+=1
        assert out is not None
    L.kth_pre_filter_batch(h, b"["+data+b"]"); L.kth_admit_queue(h, b"["+data+b","+data+b"]")
    L.kth_eval(b'{"fn":"ThrottleMetrics","throttle":'+data+b'}'); L.kth_eval(b'{"fn":"PodRequestResourceList","pod":'+data+b'}')
    if it%500==0: L.kth_metrics(h); L.kth_get_status_manifest(h,b"default",b"t")
print("calls", n, "ok")
