#!/usr/bin/env python3
"""scratch: builds one libsprb200 per combination of the K1 experiment switches into lib/variants/ and
prints the list; scripts/run_variants.py times them on the GPU (SPRB200_LIB selects the library)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sprfittingpaper2023.jl_b200")
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "lib", "variants")
os.makedirs(OUT, exist_ok=True)
BASE = dict(SPR_DEFER=0, SPR_PREFETCH=0, SPR_LDMODE=0, SPR_SCAN=0, SPR_KBIT=0)
VARIANTS = {
    "base": {},
    "ldg": dict(SPR_LDMODE=2),
    "ldg_scan": dict(SPR_LDMODE=2, SPR_SCAN=1),
    "ldg_scan_kbit": dict(SPR_LDMODE=2, SPR_SCAN=1, SPR_KBIT=1),
    "defer_ldg_scan_kbit": dict(SPR_DEFER=1, SPR_LDMODE=2, SPR_SCAN=1, SPR_KBIT=1),
}
only = sys.argv[1:]
procs = []
for name, over in VARIANTS.items():
    if only and name not in only:
        continue
    flags = dict(BASE); flags.update(over)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2",
           "--fmad=false", "-shared", "-o", os.path.join(OUT, "libsprb200_%s.so" % name)]
    cmd += ["-D%s=%d" % kv for kv in flags.items()]
    cmd += [os.path.join(CSRC, "spr_kernels.cu"), os.path.join(CSRC, "spr_api.cu"), "-lcudart"]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    out, _ = p.communicate()
    print(name, "rc", p.returncode, out[-300:] if p.returncode else "")
