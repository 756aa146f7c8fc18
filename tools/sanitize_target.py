"""Small fused + chained passes for compute-sanitizer (memcheck / racecheck / synccheck), checked against nothing: the tool is the check.
  compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kube_throttler_b200 as kt
from kube_throttler_b200 import abi, synth

for cfg, kw in (("C3", dict(m=300, n=3000, p=500)), ("C2", dict(m=200, n=2000, p=300, L=12, q_max=6)), ("C2", dict(m=40, n=70, p=33, R=1))):
    snap = synth.generate(cfg, **kw)
    eng = kt.Engine(snap.R, snap.L, snap.LN)
    eng.upload_snapshot(snap)
    eng.upload_pods_compact(abi.PODS_PENDING, abi.compact_pods(snap.pending))
    for timing in (False, True):
        eng.enable_timing(timing)
        eng.evaluate(snap.now)
        eng.sync()
    eng.download()
    eng.close()
print("sanitize target done")
