#!/usr/bin/env python
"""Top SASS instructions of the first kernel in an .ncu-rep by stall samples, with the dominant stall reasons and the
CUDA source line they belong to (needs -lineinfo and --import-source on).
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [topN]"""
import csv, io, subprocess, sys
path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows, hdr, cur_src, kernels = [], None, "", 0
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
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if r[0] != "":                     # a CUDA source line header
        cur_src = f"{cur_file}:{r[0]} {r[1].strip()[:70]}"
        continue
    try:
        n = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    stalls = {k[6:]: int(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and (v or "0").isdigit()}
    rows.append((n, d.get("Source", "").strip()[:60], cur_src, stalls, int(d.get("Instructions Executed", "0") or 0)))
tot = sum(r[0] for r in rows) or 1
agg = {}
for n, sass, src, stalls, inst in rows:
    for k, v in stalls.items():
        agg[k] = agg.get(k, 0) + v
print("total samples", tot, " stall mix:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, sass, src, stalls, inst in sorted(rows, key=lambda r: -r[0])[:topn]:
    top = ", ".join(f"{k} {v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100 * n / tot:5.2f}% inst={inst:8d}  {sass:60s} [{top}]  <- {src}")
