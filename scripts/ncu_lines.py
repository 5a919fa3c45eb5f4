#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline counters + instructions/stall samples per CUDA source line."""
import csv, subprocess, sys, io
rep = sys.argv[1]; events = float(sys.argv[2]) if len(sys.argv) > 2 else None; top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
kf = ["-k", "regex:" + sys.argv[4]] if len(sys.argv) > 4 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + kf, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.avg", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in d: print("%-80s %s %s" % (k, d[k], units[hdr.index(k)]))
for k, v in d.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v) > 0.05: print("%-80s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall "), v))
        except ValueError: pass
inst = float(d["sm__inst_executed.sum.per_cycle_elapsed"].replace(",", "")) * float(d["sm__cycles_elapsed.avg"].replace(",", ""))
print("warp-instructions executed: %.4g" % inst)
if events: print("warp-instructions per event: %.1f ; smem wavefronts per event %.1f" % (inst / events, float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",","")) / events))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kf, capture_output=True, text=True).stdout
cur = None; data = []; tot = 0; tots = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] in ("Line No", "Function Name", ""): continue
    try: ln = int(r[0]); n = int(r[7]); s = int(r[6])
    except ValueError: continue
    data.append((n, s, cur, ln, r[1])); tot += n; tots += s
data.sort(reverse=True)
print("total inst %d samples %d" % (tot, tots))
for n, s, f, ln, txt in data[:top]:
    per = (" %6.1f/ev" % (n / events)) if events else ""
    print("%6.2f%% inst%s %6.2f%% samp | %s:%d | %s" % (100 * n / tot, per, 100 * s / tots, f, ln, txt.strip()[:95]))
