#!/usr/bin/env python
"""One line per kernel launch of an ncu report: duration, warp instructions, issue-slot use, occupancy, DRAM bytes and the
six largest stall reasons.  usage: ncu_summary.py REPORT.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
num = lambda x: float(x.replace(",", "")) if x not in ("", "-") else 0.0
for r in rows[2:]:
    d = dict(zip(hdr, r))
    st = sorted(((num(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k, v in d.items()
                 if "smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k), reverse=True)[:6]
    print("%-28s %7.1f us  inst %6.2fM  issue %4.1f%%  warps %4.1f%%  lanes %4.1f  regs %s  dram r/w %5.1f/%5.1f MB  | %s" % (
        d["Kernel Name"][:28], num(d["gpu__time_duration.sum"]), num(d["smsp__inst_executed.sum"]) / 1e6,
        num(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]), num(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
        num(d["smsp__thread_inst_executed_per_inst_executed.ratio"]), d["launch__registers_per_thread"],
        num(d["dram__bytes_read.sum"]), num(d["dram__bytes_write.sum"]), " ".join("%s %.1f" % (k, v) for v, k in st)))
