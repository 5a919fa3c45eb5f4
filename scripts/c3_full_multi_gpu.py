#!/usr/bin/env python3
"""Paper-scale surrogate on all GPUs of a node (ranges of examples/cluster_scripts/
make_surrogate_metadata.jl, size (42,40,30,30,601) = 1,512,000 points), launched with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/c3_full_multi_gpu.py
One process per GPU; the whole build is ONE collective call of the library, sprb200_build_table_dist: the library
deals the points, every rank simulates its share, the library's own ncclAllGather (on the handle's stream, ordered
with its kernels: no host-side race between the gather and the scatter) assembles the slabs and a device kernel
permutes them into FirstMoment order.  torch.distributed only carries the 128-byte communicator id and the timing
reductions.  Prints the wall time (max over ranks), events, the sha256 of the table on rank 0, and re-simulates one
grid point of EVERY rank's share alone to check the assembled table bit for bit.
usage: c3_full_multi_gpu.py [n1 n2 n3 n4] [nrep]     (defaults 42 40 30 30, 100; nrep = 0 selects the reference's
default VarianceTerminator(0.01, 15, 250))"""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import torch
import torch.distributed as dist
import sprb200
from sprb200 import _lib

args = [int(a) for a in sys.argv[1:]]
size = tuple(args[:4]) if len(args) >= 4 else (42, 40, 30, 30)
nrep = args[4] if len(args) >= 5 else 100
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = sprb200.Engine(local)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if world > 1:
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(_lib.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
eng.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)
P = int(np.prod(size)); T = 601
grid = _lib.make_grid((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0), size)
tsave = np.linspace(0.0, 600.0, T)
sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 600.0, 150.0, tsave)
variance = nrep == 0
run = _lib.make_run(variance=(0.01, 15, 250), seed=20231) if variance else _lib.make_run(nsims=nrep, seed=20231)
table = torch.empty((T, P), dtype=torch.float64, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


barrier()
t0 = time.perf_counter()
_, nu, nev = eng.build_table_dist(grid, sim, run, host=False, d_table=table.data_ptr(), want_counts=True)
st = eng.stats()
t_mine = time.perf_counter() - t0
barrier()
t_total = time.perf_counter() - t0
ev = torch.tensor([float(st["events"]), float(st["trajectories"]), st["device_ms"] * 1e-3, st["table_ms"] * 1e-3,
                   st["traj_ms"] * 1e-3], dtype=torch.float64, device=dev)
mx = ev.clone()
if world > 1:
    dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    print("paper-scale surrogate %s x %s on %d GPU(s): %d points, %.0f trajectories simulated, %.4e events" %
          (size + (T,), "VarianceTerminator(0.01,15,250)" if variance else "%d replicates" % nrep, world, P, float(ev[1]), float(ev[0])))
    print("replicates entering the means: %.1f per point on average" % float(nu.mean()))
    print("build + all-gather + scatter (slowest rank, wall) %.1f s   %.3e events/s" % (t_total, float(ev[0]) / t_total))
    print("per rank simulation share: mean %.1f s, slowest %.1f s (kernel time summed over ranks: K2 %.1f s, K1 %.1f s; "
          "K1 slowest rank %.1f s)" % (float(ev[2]) / world, float(mx[2]), float(ev[3]), float(ev[4]), float(mx[4])))
    host = table.cpu().numpy()
    print("table: %.2f GB, first time slice all zero: %s, min %.4f max %.4f, sha256 %s" %
          (host.nbytes / 1e9, bool((host[0] == 0).all()), float(host.min()), float(host.max()),
           hashlib.sha256(host.tobytes()).hexdigest()))
    # integrity of the assembled table: one point of EVERY rank's share re-simulated alone (same seed, parameter
    # index = idx-1)
    rng = np.random.default_rng(3)
    ok = True
    for r in range(world):
        share = _lib.deal_indices(grid, sim, world, r, run)
        i = int(rng.choice(share))
        pt = _lib.grid_points(grid, [i])
        r1 = _lib.make_run(variance=(0.01, 15, 250), seed=20231, param_index_base=i - 1) if variance else \
            _lib.make_run(nsims=nrep, seed=20231, param_index_base=i - 1)
        out = eng.simulate(sim, pt, r1, want=("mean",))
        same = np.array_equal(out["mean"][0], host[:, i - 1])
        ok = ok and same
        print("   rank %d share, index %8d: table column %s a stand-alone run of the point" % (r, i, "==" if same else "!="))
    print("table integrity:", "OK" if ok else "MISMATCH")
eng.comm_destroy()
if world > 1:
    dist.destroy_process_group()
eng.close()
