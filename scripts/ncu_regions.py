#!/usr/bin/env python3
"""Aggregate an .ncu-rep source page into code regions of spr_kernels.cu (instructions and stall samples)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; events = float(sys.argv[2]) if len(sys.argv) > 2 else None; traj = float(sys.argv[3]) if len(sys.argv) > 3 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = collections.defaultdict(lambda: [0, 0]); tot = tots = 0
def region(f, ln):
    if f == "spr_device.cuh":
        if ln <= 31: return "philox"
        if ln <= 62: return "neg_log"
        if ln <= 80: return "rcp/dt"
        return "device.cuh other"
    if f != "spr_kernels.cu": return f
    if ln < 141: return "setup helpers (gen_position, cell_of, within_reach)"
    if ln < 198: return "setup pass 1-2 (cell sort)"
    if ln < 291: return "setup pass 3 (table)"
    if ln < 357: return "row load/apply, scan_small"
    if ln < 426: return "ticket + init state"
    if ln < 455: return "rng fetch + propensity + dt"
    if ln < 496: return "slow path (save slots)"
    if ln < 521: return "select + flip A<->B"
    if ln < 601: return "A+B->C"
    if ln < 631: return "C->A+B"
    if ln < 646: return "loop tail / final slots"
    return "other"
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] in ("Line No", "Function Name", ""): continue
    try: ln = int(r[0]); n = int(r[7]); s = int(r[6])
    except ValueError: continue
    k = region(cur, ln); agg[k][0] += n; agg[k][1] += s; tot += n; tots += s
print("region                                              inst%%   samp%%   inst/event  inst/traj")
for k, (n, s) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-50s %6.2f  %6.2f  %9.1f  %9.0f" % (k, 100 * n / tot, 100 * s / tots, n / events if events else 0, n / traj if traj else 0))
