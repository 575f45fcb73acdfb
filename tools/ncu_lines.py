#!/usr/bin/env python
"""Instructions executed and stall samples per CUDA source line of the first kernel in an .ncu-rep (needs -lineinfo and
--import-source on).   python tools/ncu_lines.py gpurun_out/x.ncu-rep [topN]"""
import csv, io, subprocess, sys
path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, rows = None, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] in ("Address", "#"):
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
if not rows:
    print(out[:2000]); sys.exit()
print(list(rows[0].keys())[:20])
agg = {}
ti = ts = 0
for d in rows:
    try:
        inst = int(d.get("Instructions Executed", "0") or 0); smp = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    op = d.get("Source", "").split()
    op = (op[1] if op and op[0].startswith("@") else op[0]) if op else "?"
    op = op.split(".")[0]
    a = agg.setdefault(op, [0, 0]); a[0] += inst; a[1] += smp
    ti += inst; ts += smp
print("total warp-inst", ti, "samples", ts)
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{k:14s} inst {i:10d} {100*i/ti:5.1f}%   samples {100*s/max(ts,1):5.1f}%")
