"""Per-role timeline of the fused pass (k_pass) from the in-kernel trace: when each role's tiles start and end relative to
the first CTA, the stage stamps of every role (median / p90 / max since pass start) and the slowest tiles of each role.
python tools/pass_trace.py [C2] [sorted|unsorted]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kube_throttler_b200 as kt
from kube_throttler_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
sort = not (len(sys.argv) > 2 and sys.argv[2] == "unsorted")
snap = synth.generate(cfg, sort_by_namespace=sort)
eng = kt.Engine(snap.R, snap.L, snap.LN)
eng.upload_snapshot(snap)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
eng.set_stream(s.cuda_stream)
eng.enable_trace(True)
STAGES = {
    "match": ["(resident: decide) rows seen", "words + pre staged", "sums seen", "constants done", "decided"],
    "reconcile": ["rows landed", "2 words evaluated", "barrier", "words+sums", "sweep"],
    "finalize": ["pre-records", "reconcile seen", "sums read", "status written"],
    "decide": ["match rows seen", "words + pre staged", "sums seen", "constants done", "decided"],
}
for it in range(4):
    with torch.cuda.stream(s):
        if it >= 2:
            flush.zero_()
            flush.sum()
        eng.evaluate(snap.now)
    torch.cuda.synchronize()
    rows, roles = eng.trace()
    rows = rows.astype(np.float64)
    t0 = rows[:, 2].min()
    lo = 0
    print(f"pass {it} ({'cold' if it >= 2 else 'warm'} L2): span {(rows[:, 3].max() - t0) / 1e3:.1f} us, {len(rows)} CTAs on {len(set(rows[:, 1]))} SMs")
    for name, n in zip(STAGES, roles):
        r = rows[lo:lo + int(n)]
        lo += int(n)
        if not len(r):
            continue
        st, en = (r[:, 2] - t0) / 1e3, (r[:, 3] - t0) / 1e3
        print(f"  {name:9s} n={int(n):4d}  start {st.min():5.1f}..{st.max():5.1f}  end med {np.median(en):5.1f} p90 {np.percentile(en, 90):5.1f} max {en.max():5.1f}  dur med {np.median(en - st):5.1f}")
        for k, label in enumerate(STAGES[name]):
            x = (r[:, 4 + k] - t0) / 1e3
            x = x[r[:, 4 + k] > 0]
            if len(x):
                print(f"            {label:18s} since pass start: med {np.median(x):5.1f}  p90 {np.percentile(x, 90):5.1f}  max {x.max():5.1f}")
        if name in ("match", "reconcile") and (r[:, 9] > 0).any():  # shared decide tiles (resident pass): the CTA's first sub-tile
            sel = r[:, 9] > 0
            print(f"            -> {int(sel.sum())} of these CTAs went on to shared decide sub-tiles")
            for k, label in enumerate(["rows + pre visible", "slots claimed", "sums seen", "constants done", "decided"]):
                x = (r[sel, 9 + k] - t0) / 1e3
                x = x[r[sel, 9 + k] > 0]
                if len(x):
                    print(f"            decide: {label:18s} since pass start: med {np.median(x):5.1f}  p90 {np.percentile(x, 90):5.1f}  max {x.max():5.1f}")
            # SM cycle counts of thread 0's way through its first sub-tile (check_decide_quad): deltas between consecutive points
            CYC = ["entry", "requests+counters", "word+claim", "ranked+pre asked", "(first dry run)", "sums visible", "pre landed", "constants", "barrier",
                   "pairs decided", "round closed"]
            c = r[sel][:, 16:27].astype(np.int64)
            dry = r[sel][:, 27]
            order = [0, 1, 2, 3, 5, 6, 7, 8, 9, 10]
            parts = []
            for a_, b_ in zip(order[:-1], order[1:]):
                okm = (c[:, a_] > 0) & (c[:, b_] > 0)
                if okm.any():
                    d = c[okm, b_] - c[okm, a_]
                    parts.append(f"{CYC[b_]} {int(np.median(d))}/{int(np.percentile(d, 90))}")
            print("            decide cycles (median/p90 since the previous point): " + " | ".join(parts))
            okm = (c[:, 4] > 0) & (c[:, 3] > 0)
            if okm.any():
                print(f"            dry runs: {int(np.median(dry))} median, {int(dry.max())} max; the first one took {int(np.median(c[okm, 4] - c[okm, 3]))} cycles (median)")
        if it == 3:
            worst = np.argsort(-en)[:4]
            for i in worst:
                stamps = " ".join(f"{(r[i, 4 + k] - t0) / 1e3:5.1f}" if r[i, 4 + k] > 0 else "    -" for k in range(len(STAGES[name])))
                print(f"            slowest: tile {i:4d} sm {int(r[i, 1]):3d} start {st[i]:5.1f} end {en[i]:5.1f} | {stamps}")
eng.close()
