#!/usr/bin/env python3
"""Per-CUDA-source-line view of an .ncu-rep (needs -lineinfo + --import-source on): share of stall samples,
of warp instructions, lanes per instruction and the long-scoreboard samples, top N lines.
usage: python profiles/lines.py <report.ncu-rep> [N]"""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows:
    if len(r) != len(hdr) or r[0] == "Line No" or r[2] != "-":  # keep the per-line totals, drop the SASS rows under them
        continue
    try:
        data.append((int(r[0]), r[1].strip()[:110], int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]),
                     int(r[ix["Thread Instructions Executed"]]), int(r[ix["stall_long_sb"]]), int(r[ix["stall_no_inst"]])))
    except ValueError:
        pass
ts, ti = sum(d[2] for d in data) or 1, sum(d[3] for d in data) or 1
print(f"samples {ts}  warp instructions {ti}")
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"{d[0]:5d} smp {100 * d[2] / ts:5.1f}% inst {100 * d[3] / ti:5.1f}% lanes {d[4] / max(d[3], 1):5.1f} long_sb {100 * d[5] / ts:5.1f}% no_inst {100 * d[6] / ts:4.1f}% | {d[1]}")
