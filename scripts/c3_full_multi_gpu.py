#!/usr/bin/env python3
"""Paper-scale surrogate on all GPUs of a node (ranges of examples/cluster_scripts/
make_surrogate_metadata.jl, size (42,40,30,30,601) = 1,512,000 points, FIXED 100 replicates), launched with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/c3_full_multi_gpu.py
Every rank builds its cost-balanced share on the device, one NCCL all-gather assembles the slabs, rank-local
scatter puts them into FirstMoment order.  Prints the wall time (max over ranks), events, an order-independent
bit checksum of the table, and re-simulates a few grid points alone to check the assembled table bit for bit.
usage: c3_full_multi_gpu.py [n1 n2 n3 n4] [nrep]     (defaults 42 40 30 30, 100; nrep = 0 selects the reference's
default VarianceTerminator(0.01, 15, 250))"""
import os, sys, time
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
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
P = int(np.prod(size)); T = 601
grid = _lib.make_grid((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0), size)
tsave = np.linspace(0.0, 600.0, T)
sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 600.0, 150.0, tsave)
variance = nrep == 0
run = _lib.make_run(variance=(0.01, 15, 250), seed=20231) if variance else _lib.make_run(nsims=nrep, seed=20231)
t0 = time.perf_counter()
cost = _lib.estimate_cost(sim, _lib.grid_points(grid))
deal = [d + 1 for d in _lib.balanced_deal(cost, world)]
# all_gather needs equal slabs: pad the short ranks by repeating their last index (the scatter just rewrites it)
count = max(len(d) for d in deal)
deal = [np.concatenate([d, np.repeat(d[-1:], count - len(d))]) for d in deal]
idx = deal[rank]; idx_all = np.concatenate(deal)
t_deal = time.perf_counter() - t0
eng = sprb200.Engine(local)
rows = torch.empty((count, T), dtype=torch.float64, device=dev)
nev = torch.zeros(count, dtype=torch.int64, device=dev)
nused = torch.zeros(count, dtype=torch.int32, device=dev)
all_rows = torch.empty((count * world, T), dtype=torch.float64, device=dev) if world > 1 else rows
table = torch.empty((T, P), dtype=torch.float64, device=dev)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
barrier()
t0 = time.perf_counter()
eng.build_indices_device(grid, sim, run, idx, rows.data_ptr(), nused.data_ptr(), nev.data_ptr())
st = eng.stats()
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
barrier()
t_build_all = time.perf_counter() - t0
if world > 1:
    dist.all_gather_into_tensor(all_rows, rows)
eng.scatter_rows_device(all_rows.data_ptr(), idx_all, T, P, table.data_ptr())
barrier()
t_total = time.perf_counter() - t0
ev = torch.tensor([float(st["events"]), float(st["trajectories"]), t_build, st["table_ms"] * 1e-3, st["traj_ms"] * 1e-3,
                   float(nused[:len(idx)].double().sum())], dtype=torch.float64, device=dev)
mx = ev.clone()
if world > 1:
    dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
checksum = int(table.view(torch.int64).sum().item())
if rank == 0:
    print("paper-scale surrogate %s x %s on %d GPU(s): %d points, %.0f trajectories simulated, %.4e events" %
          (size + (T,), "VarianceTerminator(0.01,15,250)" if variance else "%d replicates" % nrep, world, P, float(ev[1]), float(ev[0])))
    print("replicates entering the means: %.1f per point on average" % (float(ev[5]) / (count * world)))
    print("build (slowest rank) %.1f s   build + all-gather + scatter %.1f s   (host-side deal %.1f s)   %.3e events/s" %
          (t_build_all, t_total, t_deal, float(ev[0]) / t_build_all))
    print("per rank: build mean %.1f s, slowest %.1f s (kernel time summed over ranks: K2 %.1f s, K1 %.1f s)" %
          (float(ev[2]) / world, float(mx[2]), float(ev[3]), float(ev[4])))
    print("table: %.2f GB, first time slice all zero: %s, min %.4f max %.4f, int64 bit-sum checksum %d" %
          (table.numel() * 8 / 1e9, bool((table[0] == 0).all().item()), float(table.min()), float(table.max()), checksum))
    # integrity of the assembled table: re-simulate a few points alone (same seed, parameter index = idx-1)
    rng = np.random.default_rng(3)
    ok = True
    for i in rng.choice(P, size=5, replace=False) + 1:
        pt = _lib.grid_points(grid, [i])
        r1 = _lib.make_run(variance=(0.01, 15, 250), seed=20231, param_index_base=int(i) - 1) if variance else \
            _lib.make_run(nsims=nrep, seed=20231, param_index_base=int(i) - 1)
        out = eng.simulate(sim, pt, r1, want=("mean",))
        same = np.array_equal(out["mean"][0], table[:, int(i) - 1].cpu().numpy())
        ok = ok and same
        print("   index %8d: table column %s a stand-alone run of the point" % (i, "==" if same else "!="))
    print("table integrity:", "OK" if ok else "MISMATCH")
if world > 1:
    dist.destroy_process_group()
eng.close()
