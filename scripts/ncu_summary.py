#!/usr/bin/env python3
"""profiles/r2_ncu_k1.json from an ncu --set full capture of ONE spr_traj_kernel launch: the counters bench.py
quotes as roofline.measured (executed issue-slot fraction, executed warp-instructions per event, DRAM bytes of
the launch).  usage: ncu_summary.py <file.ncu-rep> <events in that launch> <out.json> [note]"""
import csv, io, json, subprocess, sys
rep, events, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:spr_traj_kernel"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
f = lambda k: float(d[k].replace(",", ""))
def bytes_of(k):
    v, un = f(k), u[k].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[un]
inst = f("sm__inst_executed.sum.per_cycle_elapsed") * f("sm__cycles_elapsed.avg")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
          for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v not in ("", "n/a") and float(v) > 0.05}
res = {
    "kernel": "spr_traj_kernel", "source": rep.split("/")[-1], "commit": commit, "note": note,
    "events_in_launch": events, "launch_ms": f("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3}[u["gpu__time_duration.sum"]],
    "registers_per_thread": f("launch__registers_per_thread"), "grid": f("launch__grid_size"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "ipc_per_sm": f("sm__inst_executed.avg.per_cycle_active"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "warp_latency_per_inst_cycles": f("smsp__average_warp_latency_per_inst_issued.ratio"),
    "executed_warp_inst_per_event": inst / events,
    "smem_wavefronts_per_event": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / events,
    "smem_pipe_pct": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    "l1_hit_rate_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_rate_pct": f("lts__t_sector_hit_rate.pct"),
    "dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
    "dram_bytes_read": bytes_of("dram__bytes_read.sum"), "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
    "stall_cycles_per_issue": stalls,
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
