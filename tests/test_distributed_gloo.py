"""CPU, world_size = 2, gloo: the strip/halo communication logic of more3d_b200.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from more3d_b200.distributed import exchange_halos, allreduce_minmax, strip_bounds
    try:
        rng = np.random.Generator(np.random.PCG64(100))
        pts = rng.random((1000, 3)) * [10.0, 4.0, 1.0]
        edges = strip_bounds(0.0, 10.0, world)
        lo, hi = edges[rank], edges[rank + 1]
        halo = 0.7
        own = pts[(pts[:, 0] >= lo) & ((pts[:, 0] < hi) | (rank == world - 1))]
        own_t = torch.from_numpy(own)
        left = own_t[own_t[:, 0] < lo + halo]
        right = own_t[own_t[:, 0] >= hi - halo]
        fl, fr = exchange_halos(left, right, rank, world)
        # what the neighbours should have sent
        want_l = pts[(pts[:, 0] >= lo - halo) & (pts[:, 0] < lo)] if rank > 0 else np.zeros((0, 3))
        want_r = pts[(pts[:, 0] >= hi) & (pts[:, 0] < hi + halo)] if rank < world - 1 else np.zeros((0, 3))
        ok = np.array_equal(fl.numpy(), want_l) and np.array_equal(fr.numpy(), want_r)
        mm = torch.tensor([1.0 + rank, 5.0 - rank, -2.0 * rank, 3.0 * rank], dtype=torch.float32)
        allreduce_minmax(mm)
        ok = ok and mm.tolist() == [1.0, 5.0, -2.0 * (world - 1), 3.0 * (world - 1)]
        # sharded Otsu: all-reduced range + histogram == Otsu of the whole column (oracle restatement)
        from more3d_b200.distributed import global_otsu
        from oracle import ref_path as rp
        col = (np.random.Generator(np.random.PCG64(7)).random(5000) ** 3).astype(np.float32)
        mine = torch.from_numpy(col[rank::world].copy())

        def np_hist(v, edges):
            return torch.from_numpy(np.histogram(v.numpy().astype(np.float64), bins=edges)[0].astype(np.int64))
        ok = ok and global_otsu(mine, np_hist) == rp.otsu_threshold(col)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_halo_exchange_and_minmax_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_strip_bounds():
    from more3d_b200.distributed import strip_bounds
    e = strip_bounds(-4.0, 4.0, 4)
    assert np.allclose(e, [-4, -2, 0, 2, 4])
