#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an .ncu-rep source page.
usage: python tools/ncu_hot.py rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][ix["# Samples"]]))[:n]:
    s = int(r[ix["# Samples"]])
    top = sorted(((int(r[ix[h]]), h) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {s:7d} {100*s/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {top}")
