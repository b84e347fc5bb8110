#!/usr/bin/env python
"""Hot SASS instructions of an .ncu-rep source page: samples, executed count, top stall reason.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top_n] [--all]"""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
st_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "instructions", inst)
agg = {}
for r in data:
    for h in st_cols:
        agg[h] = agg.get(h, 0) + int(r[ci[h]] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
if "--all" in sys.argv:
    for n, r in enumerate(data):
        s = int(r[ci["# Samples"]] or 0)
        st = max(st_cols, key=lambda h: int(r[ci[h]] or 0))
        print(f"{n:5d} {s:7d} {int(r[ci['Instructions Executed']] or 0):9d}  {r[ci['Source']].strip()[:90]:90s} {st if s else ''}")
else:
    order = sorted(range(len(data)), key=lambda n: -int(data[n][ci["# Samples"]] or 0))[:top]
    for n in sorted(order):
        r = data[n]; s = int(r[ci["# Samples"]] or 0)
        st = sorted(((int(r[ci[h]] or 0), h) for h in st_cols), reverse=True)[:2]
        print(f"{n:5d} {s:7d} {100*s/tot:5.1f}% {int(r[ci['Instructions Executed']] or 0):9d}  {r[ci['Source']].strip()[:80]:80s} {st}")
