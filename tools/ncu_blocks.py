"""Instruction counts of an .ncu-rep by straight-line run of equal execution count (≈ basic blocks)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; thresh = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 1:] if len(r) == len(hdr)]
runs = []
for r in data:
    c = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    if runs and runs[-1][0] == c: runs[-1][1] += 1; runs[-1][2] += s
    else: runs.append([c, 1, s, r[ix["Source"]][:60]])
tot = sum(c * n for c, n, _, _ in runs)
print("total warp-instructions", tot, "static", len(data))
for c, n, s, first in runs:
    if c * n > thresh: print(f"exec x{c:7d}  n_instr {n:4d}  total {c*n:9d} ({100*c*n/tot:4.1f}%)  samples {s:5d}  first: {first}")
