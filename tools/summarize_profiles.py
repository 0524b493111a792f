#!/usr/bin/env python
"""Turn gpurun_out/<tag>_launches.csv (+ <tag>_top.ncu-rep) into the tracked summaries under profiles/ (usage: summarize_profiles.py r02)."""
import csv, collections, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
lines = [l for l in open(os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
def us(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
names = [r["Kernel Name"] for r in rows]
flush = [i for i, n in enumerate(names) if "FillFunctor<unsigned char>" in n]
# steps of the first timed() region (the C-ABI session path): between consecutive L2 flushes
steps = []
for a, b in zip(flush, flush[1:]):
    seg = rows[a + 1:b]
    if any("dibr_forward" in r["Kernel Name"] for r in seg) and len(seg) < 40:
        steps.append(seg)
seg = steps[1]
agg = collections.OrderedDict(); tot = 0.0
for r in seg:
    n = r["Kernel Name"].split("(")[0][:70]; agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += us(r); tot += us(r)
with open(os.path.join(out_dir, f"{tag}_launches_step.csv"), "w") as f:
    w = csv.writer(f); w.writerow(["kernel", "launches_per_step", "us_per_step", "share"])
    for n, (c, t) in agg.items(): w.writerow([n, c, f"{t:.1f}", f"{t / tot:.3f}"])
    w.writerow(["TOTAL", sum(c for c, _ in agg.values()), f"{tot:.1f}", "1.000"])
# full launch list, trimmed to the columns that matter
with open(os.path.join(out_dir, f"{tag}_launches_full.csv"), "w") as f:
    w = csv.writer(f); w.writerow(["id", "kernel", "grid", "block", "duration_us"])
    for r in rows: w.writerow([r["ID"], r["Kernel Name"][:90], r["Grid Size"], r["Block Size"], f"{us(r):.2f}"])
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_top.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw))); hdr, units = rr[0], rr[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
            "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(out_dir, f"{tag}_ncu_full_metrics.csv"), "w") as f:
        w = csv.writer(f); w.writerow(["metric", "unit"] + [r[idx["Kernel Name"]][:40] for r in rr[2:]])
        for m in want[1:]:
            if m in idx: w.writerow([m, units[idx[m]]] + [r[idx[m]] for r in rr[2:]])
    # traffic per launch of the dominant kernel for bench.py's roofline.traffic
    for r in rr[2:]:
        if "dibr_forward" in r[idx["Kernel Name"]]:
            def mb(x, u): 
                v = float(x.replace(",", "")); return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[u]
            t = mb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + mb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            def num(m):
                return float(r[idx[m]].replace(",", "")) if m in idx and r[idx[m]] not in ("", "n/a") else None
            sys.path.insert(0, ROOT)
            import bench
            json.dump({"csrc_hash": bench.csrc_hash(),        # of the tree the capture was made on (run this right after the capture)
                       "dibr_forward_kernel_bytes_per_launch": t,
                       "warp_instructions_per_launch": num("smsp__inst_executed.sum"),
                       "issue_slots_busy_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "warp_slots_occupied_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                       "active_lanes_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
                       "source": f"profiles/{tag}_ncu_full_metrics.csv (ncu --set full, student pass)"},
                      open(os.path.join(out_dir, "traffic.json"), "w"))
            break
print(open(os.path.join(out_dir, f"{tag}_launches_step.csv")).read())
