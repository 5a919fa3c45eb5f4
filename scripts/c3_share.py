"""Paper-scale surrogate (examples/cluster_scripts/make_surrogate_metadata.jl ranges, size
(42,40,30,30,601) = 1,512,000 points): the share of ONE rank of `world` ranks (block-cyclic over the linear
index, block = 1680 points = one (konb, reach) pair), FIXED 100 replicates -- what each GPU of an 8-GPU
build does.  Prints time, events and a checksum; tests/tools/c3_sample_parity.py checks random points of the grid against
the CPU oracle of the reference algorithm.
usage: c3_share.py [world=8] [rank=0] [nrep=100] [variance]"""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import sprb200
from sprb200 import _lib
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nrep = int(sys.argv[3]) if len(sys.argv) > 3 else 100
variance = len(sys.argv) > 4 and sys.argv[4] == "variance"
size = (42, 40, 30, 30)
P = int(np.prod(size))
grid = _lib.make_grid((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0), size)
blk = size[0] * size[1]
nblocks = P // blk
idx = np.concatenate([np.arange(b * blk + 1, (b + 1) * blk + 1, dtype=np.int64) for b in range(rank, nblocks, world)])
eng = sprb200.Engine(0)
L = _lib.domain_length()
tsave = np.linspace(0.0, 600.0, 601)
sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, tsave)
run = _lib.make_run(variance=(0.01, 15, 250), seed=20231) if variance else _lib.make_run(nsims=nrep, seed=20231)
t0 = time.perf_counter()
rows, nu, nev = eng.build_indices(grid, sim, run, idx)
wall = time.perf_counter() - t0
st = eng.stats()
print("C3 share: rank %d of %d, %d of %d points, %s" % (rank, world, len(idx), P, "VarianceTerminator(.01,15,250)" if variance else "FIXED %d" % nrep))
print("wall %.1f s  device %.1f s (K2 %.1f, K1 %.1f)  trajectories %d  events %.4e  ev/s %.3e  launches %d" %
      (wall, st["device_ms"] * 1e-3, st["table_ms"] * 1e-3, st["traj_ms"] * 1e-3, st["trajectories"], st["events"],
       st["events"] / (st["device_ms"] * 1e-3), st["kernel_launches"]), flush=True)
print("n_used mean %.1f min %d max %d; events/trajectory mean %.0f max-point %.0f; rows sha256 %s" %
      (nu.mean(), nu.min(), nu.max(), nev.sum() / max(st["trajectories"], 1), (nev / np.maximum(nu, 1)).max(),
       hashlib.sha256(rows.tobytes()).hexdigest()[:16]))
assert np.all(rows[:, 0] == 0.0) and np.all(rows >= 0.0) and np.all(rows <= 1.0)
