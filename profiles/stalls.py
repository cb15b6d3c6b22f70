#!/usr/bin/env python3
"""Stall-reason totals of an .ncu-rep (warp sampling over the whole kernel). usage: python profiles/stalls.py <rep>"""
import csv, io, subprocess, sys
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0].startswith("#") is False and "# Samples" in r)
ix = {h: i for i, h in enumerate(hdr)}
cols = [h for h in hdr if h.startswith("stall_")]
tot = {c: 0 for c in cols}
for r in rows:
    if len(r) != len(hdr) or r is hdr:
        continue
    for c in cols:
        try:
            tot[c] += int(r[ix[c]])
        except ValueError:
            pass
s = sum(tot.values()) or 1
for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    print(f"{c:28s} {100 * v / s:5.1f}%")
