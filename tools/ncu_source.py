#!/usr/bin/env python
"""Per-source-line stall-sample summary of the FIRST kernel in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_source.py gpurun_out/x.ncu-rep [topN]"""
import collections
import csv
import io
import subprocess
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur_file, hdr, kernels = None, None, 0
agg = collections.OrderedDict()
cur_key = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "Function Name":
        continue
    d = dict(zip(hdr, r))
    if r[0] != "":
        try:
            cur_key = (cur_file, int(r[0]), r[1].strip()[:100])
        except ValueError:
            continue
        agg.setdefault(cur_key, [0, 0])
        continue
    if cur_key is None:
        continue
    try:
        agg[cur_key][0] += int(d.get("# Samples", "0") or 0)
        agg[cur_key][1] += int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
print(f"total samples {tot}")
for (f, line, src), (s, inst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100 * s / tot:6.2f}%  inst={inst:9d}  {f}:{line:<4d} {src}")
