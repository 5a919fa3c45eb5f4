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


GRID_SIZE = (5, 3, 3, 3)          # 135 points: not divisible by 2 or 4, so the padded-slab path is exercised


def _job():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    import sprb200
    from sprb200 import _lib
    grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), GRID_SIZE)
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 600.0, 150.0, np.linspace(0, 600, 601))
    return _lib, grid, sim


def _worker(rank, world, port, tmp):
    _lib, grid, sim = _job()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 7
    # what sprb200_build_table_dist does on every rank: the library's deal, a slab padded to the largest share,
    # one all-gather, a scatter that skips the padding rows (index 0)
    deal = [_lib.deal_indices(grid, sim, world, r) for r in range(world)]
    maxcount = max(len(d) for d in deal)
    idx = deal[rank]
    slab = np.zeros((maxcount, T))
    slab[:len(idx)] = fake_rows(idx, T)
    rows = torch.from_numpy(slab)
    gathered = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(gathered, rows)
    all_rows = torch.cat(gathered).numpy()
    idx_all = np.zeros(maxcount * world, dtype=np.int64)
    for r in range(world):
        idx_all[r * maxcount:r * maxcount + len(deal[r])] = deal[r]
    P = int(np.prod(GRID_SIZE))
    table = np.zeros((T, P))
    for k, i in enumerate(idx_all):
        if i >= 1:     # sprb200_merge_slice: table[s*P + idx-1] = rows[k*T + s], the device scatter's arithmetic
            _lib.merge_slice(grid, T, all_rows[k:k + 1], int(i), int(i), table)
    np.save(os.path.join(tmp, "table_%d.npy" % rank), table)
    dist.barrier()
    dist.destroy_process_group()


def test_rank_indices_partition(built_lib):
    _lib, grid, sim = _job()
    P = int(np.prod(GRID_SIZE))
    for world in (1, 2, 4, 8):
        deal = [_lib.deal_indices(grid, sim, world, r) for r in range(world)]
        allidx = np.concatenate(deal)
        assert len(allidx) == P and np.array_equal(np.sort(allidx), np.arange(1, P + 1))
        sizes = [len(d) for d in deal]
        assert max(sizes) - min(sizes) <= 1
        assert all(np.all(np.diff(d) > 0) for d in deal)


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
