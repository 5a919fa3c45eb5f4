#!/usr/bin/env python3
"""Static evidence from the built library (no GPU needed): registers / spills / barriers per kernel from
`ptxas -v` (lib/ptxas_info.txt, written by build.py) and, from `cuobjdump -sass`, the SASS instruction count of every
kernel with the tensor-memory instructions (tcgen05.alloc/dealloc/relinquish = UTCATOMSWS, tcgen05.ld = LDTM,
tcgen05.st = STTM), convergence checks (BRA.DIV) and BSSY/BSYNC pairs.
usage: python scripts/sass_summary.py > profiles/r2_sass_ptxas_summary.txt"""
import os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sprfittingpaper2023.jl_b200", "lib")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [o.replace("spr::(anonymous namespace)::", "").replace("unsigned short", "u16").replace("unsigned int", "u32")
            .replace("(spr::TrajArgs)", "").replace("(spr::TableArgs)", "").replace("(spr::FitArgs)", "").replace("void ", "")
            for o in out]


regs = {}
cur = None
for line in open(os.path.join(LIB, "ptxas_info.txt")):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1); regs[cur] = {}
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur and "spill" not in regs[cur]:
        regs[cur]["spill"] = "%s/%s" % (m.group(2), m.group(3))
    m = re.search(r"Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes smem)?", line)
    if m and cur:
        regs[cur].update(reg=m.group(1), bar=m.group(2), smem=m.group(3) or "0")
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, "libsprb200.so")], capture_output=True, text=True).stdout
kern, name = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); kern[name] = dict(n=0, LDTM=0, STTM=0, UTCATOMSWS=0, DIV=0, BSSY=0, LDS=0, STS=0, LDG=0, SHFL=0, DFMA=0, MUFU=0)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1); k = kern[name]; k["n"] += 1
        base = op.split(".")[0]
        if base in k:
            k[base] += 1
        if op.startswith("BRA.DIV"):
            k["DIV"] += 1
names = list(kern)
print("# python scripts/sass_summary.py  (ptxas -v of build.py + cuobjdump -sass of lib/libsprb200.so; CUDA 12.9, sm_100a)")
print("# spill = bytes spill stores/loads; TM = tensor-memory instructions: UTCATOMSWS (tcgen05.alloc / dealloc / relinquish_alloc_permit),")
print("# LDTM (tcgen05.ld), STTM (tcgen05.st); DIV = BRA.DIV convergence checks; BSSY = reconvergence regions")
print("%-58s %4s %7s %4s %6s | %5s %4s %4s %4s %4s %4s %4s %4s %4s %4s %4s" %
      ("kernel", "regs", "spill", "bar", "smem", "SASS", "UTCA", "LDTM", "STTM", "DIV", "BSSY", "LDS", "STS", "LDG", "SHFL", "DFMA"))
for nm, dm in zip(names, demangle(names)):
    r, k = regs.get(nm, {}), kern[nm]
    print("%-58s %4s %7s %4s %6s | %5d %4d %4d %4d %4d %4d %4d %4d %4d %4d %4d" %
          (dm[:58], r.get("reg", "?"), r.get("spill", "?"), r.get("bar", "?"), r.get("smem", "?"), k["n"], k["UTCATOMSWS"],
           k["LDTM"], k["STTM"], k["DIV"], k["BSSY"], k["LDS"], k["STS"], k["LDG"], k["SHFL"], k["DFMA"]))
