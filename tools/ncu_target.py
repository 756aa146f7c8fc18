"""Small driver for ncu: uploads a BASELINE config and runs a few passes (no torch, no oracle).
  ncu --set full --clock-control none --import-source on -k regex:k_reconcile -s 2 -c 1 -o gpurun_out/prof python tools/ncu_target.py C2 5
  ncu ... -k regex:k_pass ... python tools/ncu_target.py C2 5 fused     (the one-launch pass; the default runs the three chained kernels)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kube_throttler_b200 as kt
from kube_throttler_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
scale = int(os.environ.get("KT_ROWS_SCALE", "1"))
if scale > 1:
    base = synth.CONFIGS[cfg]
    snap = synth.generate(cfg, n=base["n"] * scale, p=base["p"] * scale, calibrate=False)
else:
    snap = synth.generate(cfg, calibrate=cfg in ("C1", "C2"))
eng = kt.Engine(snap.R, snap.L, snap.LN)
eng.upload_snapshot(snap)
fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
eng.enable_timing(not fused)
for i in range(iters):
    eng.evaluate(snap.now)
    eng.sync()
    t = eng.timing()
    if fused:
        print(f"pass {i}: fused, {t.launches} launch")
        continue
    print(f"pass {i}: reconcile {t.reconcile_ms*1e3:.1f} us  finalize {t.finalize_ms*1e3:.1f} us  check {t.check_ms*1e3:.1f} us  total {t.total_ms*1e3:.1f} us")
eng.close()
