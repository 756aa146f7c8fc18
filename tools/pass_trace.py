"""Per-role timeline of the fused pass (k_pass) from the in-kernel trace: when each role's tiles start and end
relative to the first CTA.  python tools/pass_trace.py [C2]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kube_throttler_b200 as kt
from kube_throttler_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
snap = synth.generate(cfg, calibrate=cfg in ("C1", "C2"))
eng = kt.Engine(snap.R, snap.L, snap.LN)
eng.upload_snapshot(snap)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
eng.set_stream(s.cuda_stream)
eng.enable_trace(True)
for it in range(4):
    with torch.cuda.stream(s):
        if it >= 2:
            flush.zero_()
            flush.sum()
        eng.evaluate(snap.now)
    torch.cuda.synchronize()
    rows, roles = eng.trace()
    t0 = rows[:, 2].min()
    names = ["match", "reconcile", "finalize", "decide"]
    lo = 0
    print(f"pass {it} ({'cold' if it >= 2 else 'warm'} L2): span {(rows[:, 3].max() - t0) / 1e3:.1f} us, {len(rows)} CTAs on {len(set(rows[:, 1]))} SMs")
    for name, n in zip(names, roles):
        r = rows[lo:lo + int(n)]
        lo += int(n)
        st, en = (r[:, 2] - t0) / 1e3, (r[:, 3] - t0) / 1e3
        print(f"  {name:9s} n={int(n):4d}  start {st.min():6.1f}..{st.max():6.1f}  end {en.min():6.1f}..{en.max():6.1f}  median dur {np.median(en - st):6.1f} us")
        if name == "finalize":
            stages = [(r[:, k].astype(np.float64) - t0) / 1e3 for k in (4, 5, 6, 7)]
            print("            finalize stages since pass start (median / max us): pre-records %.1f/%.1f | reconcile seen %.1f/%.1f | sums read %.1f/%.1f | status written %.1f/%.1f"
                  % tuple(v for x in stages for v in (np.median(x), x.max())))
        if name == "reconcile":
            stages = [(r[:, k].astype(np.float64) - r[:, 2].astype(np.float64)) / 1e3 for k in (4, 5, 6, 7)]
            print("            reconcile stages since tile start (median us): rows loaded+translated %.1f | barrier %.1f | words+sums %.1f | sweep %.1f"
                  % tuple(np.median(x) for x in stages))
eng.close()
