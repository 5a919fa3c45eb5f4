"""CPU: the cell-list row builder of K2 (`scan_range` / `particle_row` in csrc/spr_kernels.cu) compiled for the host.

The device functions are extracted verbatim from the .cu file (tests/c/k2_rows_harness.cpp supplies `uint4`,
`__umulhi`, `__ffs`) and run against an O(N^2) scan of the same lattice positions: DIM 2 and 3, reach from 0 to
beyond the box (all-pairs path), coincident particles and particles on the box faces, staging rows that overflow.
Every build variant (SPRB200_K2_SKIP 0 = all 3^DIM cells / 1 = per-axis cell skipping, with and without the signed
squares and the two-candidate loop) must give the same neighbours in the same order: skipping is an optimisation, not a change of the specification
(forward_simulator.jl:37-61: `periodic_dist_sq <= reach^2`)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = os.path.join(ROOT, "sprfittingpaper2023.jl_b200", "csrc", "spr_kernels.cu")


def _extract(dst):
    s = open(KERNELS).read()
    k2 = s[s.index("struct K2Smem {"):s.index("__host__ __device__ inline int k2_align16")]
    body = s[s.index("struct RowPacker {"):s.index("template <int DIM>\n__device__ __forceinline__ uint32_t k2_cell")]
    with open(os.path.join(dst, "extract.inc"), "w") as f:
        f.write(k2 + "\n" + body)


def test_particle_rows_match_all_pairs_scan_in_every_variant(tmp_path):
    d = str(tmp_path)
    _extract(d)
    outs = []
    for n, flags in enumerate((("-DSPRB200_K2_SKIP=0", "-DSPRB200_K2_SIGNED=0", "-DSPRB200_K2_UNROLL2=0"),
                               ("-DSPRB200_K2_SKIP=1", "-DSPRB200_K2_SIGNED=0", "-DSPRB200_K2_UNROLL2=1"),
                               ("-DSPRB200_K2_SKIP=1", "-DSPRB200_K2_SIGNED=1", "-DSPRB200_K2_UNROLL2=1"))):
        exe = os.path.join(d, "h%d" % n)
        subprocess.run(["g++", "-O1", "-std=c++17", "-I", d, *flags, "-o", exe,
                        os.path.join(ROOT, "tests", "c", "k2_rows_harness.cpp")], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "TOTAL bad 0" in r.stderr
        outs.append(r.stdout)
    assert len(outs[0]) > 100000                      # the ordered edge lists
    assert outs[0] == outs[1] == outs[2]
