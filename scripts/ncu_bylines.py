#!/usr/bin/env python3
"""All source lines of an .ncu-rep in file/line order with warp-instructions per event and stall-sample share."""
import csv, subprocess, sys, io
rep = sys.argv[1]; events = float(sys.argv[2]); fname = sys.argv[3] if len(sys.argv) > 3 else "spr_kernels.cu"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; data = {}; tot = tots = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] in ("Line No", "Function Name", ""): continue
    try: ln = int(r[0]); n = int(r[7]); s = int(r[6])
    except ValueError: continue
    tot += n; tots += s
    if n or s:
        k = (cur, ln); v = data.get(k, [0, 0, r[1]]); v[0] += n; v[1] += s; data[k] = v
print("total %.1f inst/event" % (tot / events))
for (f, ln), (n, s, txt) in sorted(data.items()):
    if f.endswith(".hpp") or n / events >= 0.05:
        print("%-18s %4d %7.2f/ev %5.2f%%s | %s" % (f[:18], ln, n / events, 100 * s / tots, txt.strip()[:100]))
