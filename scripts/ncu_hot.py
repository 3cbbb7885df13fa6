#!/usr/bin/env python
"""Top-sampled SASS instructions of an .ncu-rep (no GPU needed), with the dominant stall columns.
usage: python scripts/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = rows[hi[0] + 1:end]
cs, sc, ci = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[cs]) for r in body if r[cs].isdigit())
idx = sorted(range(len(body)), key=lambda i: -int(body[i][cs] or 0))[:N]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100*int(r[cs])/tot:5.2f}% exec={int(r[ci]):>9d} {r[sc][:70]:70s} " + ' '.join(f"{n[6:]}={v}" for v, n in st if v))
