"""CPU, world_size 2 over gloo: the multi-rank plumbing of the surrogate build (cost-balanced index
deal, all-gather of the per-rank slabs, scatter into FirstMoment order) without GPUs.  The per-rank
'simulation' is a deterministic function of the parameter index, exactly the property the GPU path has
(RNG keyed by parameter index and replicate), so the gathered table must equal the single-rank one."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_rows(idx, T):
    """stand-in for build_indices: row k depends only on the global index"""
    s = np.arange(T, dtype=np.float64)[None, :]
    return np.sin(0.001 * idx[:, None].astype(np.float64) * (s + 1.0))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    import bench
    from sprb200 import _lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 7
    idx = bench.rank_indices(world, rank)
    rows = torch.from_numpy(fake_rows(idx, T))
    gathered = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(gathered, rows)
    all_rows = torch.cat(gathered).numpy()
    idx_all = np.concatenate([bench.rank_indices(world, r) for r in range(world)])
    P = len(idx_all)
    grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), bench.NK + (bench.NREACH_PER_GPU * world,))
    table = np.zeros((T, P))
    # host scatter = sprb200_merge_slice per run of consecutive indices (same arithmetic as the device scatter
    # sprb200_scatter_rows_device: table[s*P + idx-1] = rows[k*T + s])
    cuts = np.flatnonzero(np.diff(idx_all) != 1) + 1
    for a, b in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [len(idx_all)]])):
        _lib.merge_slice(grid, T, all_rows[a:b], int(idx_all[a]), int(idx_all[b - 1]), table)
    np.save(os.path.join(tmp, "table_%d.npy" % rank), table)
    dist.barrier()
    dist.destroy_process_group()


def test_rank_indices_partition():
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8):
        allidx = np.concatenate([bench.rank_indices(world, r) for r in range(world)])
        P = bench.NK[0] * bench.NK[1] * bench.NK[2] * bench.NREACH_PER_GPU * world
        assert len(allidx) == P and np.array_equal(np.sort(allidx), np.arange(1, P + 1))
        sizes = {len(bench.rank_indices(world, r)) for r in range(world)}
        assert sizes == {P // world}                       # equal slabs (all_gather needs them)


def test_two_rank_gather_equals_single_rank(built_lib, tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    t0 = np.load(tmp_path / "table_0.npy")
    t1 = np.load(tmp_path / "table_1.npy")
    assert np.array_equal(t0, t1)
    P = t0.shape[1]
    expect = fake_rows(np.arange(1, P + 1), 7).T
    assert np.array_equal(t0, expect)
