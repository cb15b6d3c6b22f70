#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key counters, opcode mix and stall reasons per read.
usage: python profiles/analyze.py <report.ncu-rep> <reads in the profiled launch> [out.json [workload]]
With a workload name the facts bench.py reads live (DRAM bytes per launch, warp instructions per read) are also written
to profiles/traffic.json under that name."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

rep, n_reads = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
m = dict(zip(rows[0], rows[2]))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
out = {k: m.get(k) for k in keys}
units = dict(zip(rows[0], rows[1]))
out["units"] = {k: units.get(k) for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum")}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hdr = srows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stalls = collections.Counter(); total = 0
scols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in srows[2:]:
    try:
        n = int(r[ix["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op = ".".join((mm.group(2) if mm else "?").split(".")[:2])
    ops[op] += n; total += n
    for c in scols:
        try:
            stalls[c] += int(r[ix[c]])
        except ValueError:
            pass
wr = n_reads / 32
out["reads_in_launch"] = n_reads
out["warp_instructions_per_read"] = total / n_reads  # warp-level instructions issued per read (a warp serves 32 reads)
out["instructions_per_read"] = total / wr            # the same per thread: what one read costs in instructions
out["opcode_mix_per_read"] = {k: round(v / wr, 2) for k, v in ops.most_common(24)}
ts = sum(stalls.values()) or 1
out["stall_pct"] = {k: round(100 * v / ts, 1) for k, v in stalls.most_common(8)}
print(json.dumps(out, indent=1))
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], "w"), indent=1)
if len(sys.argv) > 4:
    import os
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    tj = json.load(open(tp)) if os.path.exists(tp) else {}
    tj.setdefault("workloads", {})[sys.argv[4]] = {
        "summary": "profiles/" + os.path.basename(sys.argv[3]), "reads_in_launch": n_reads,
        "dram_bytes_per_launch": sum(float(out[k] or 0) * scale.get(out["units"].get(k), 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")),
        "warp_inst_per_read": float(out["smsp__inst_executed.sum"]) / n_reads, "thread_inst_per_read": out["instructions_per_read"],
        "issue_active_pct": float(out["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "fmaheavy_pct": float(out["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
        "kernel_ms_under_ncu": float(out["gpu__time_duration.sum"])}
    json.dump(tj, open(tp, "w"), indent=1)
