"""GPU, real multi-rank: 2 processes over NCCL (one per GPU) against the untiled single-GPU result.
Skipped on boxes with fewer than two GPUs (`gpurun --gpus 2` runs it); the logical-rank tests in test_gpu_strips.py
cover the same arithmetic on one GPU and tests/test_distributed_gloo.py the communication logic on CPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RADII = (0.5, 0.75, 1.0)
# rough tile (noise 0.15 m) + sigma_normal 1e-3: on the smooth default relief every dn is zeroed, and a constant column
# normalises to NaN (max == min in DM_saliency.py:577-578) -- the comparison would be vacuous
PARAMS = dict(roughness=0.02, alpha=0.05, sigma_normal=1e-3, sigma_curvature=0.01, min_points=4)
TOTAL = 600_000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from more3d_b200 import synthetic
        from more3d_b200.distributed import StripSaliency
        per = TOTAL // world
        lx, _ = synthetic.hash_tile_extent(TOTAL)
        own = synthetic.hash_terrain_device(11, rank * per, per, TOTAL, noise=0.15, device=dev)
        x_lo = rank * lx / world - 0.5 * lx
        job = StripSaliency(per, rank, world, dev, RADII, PARAMS, 0.5, cloud=own, x_range=(x_lo, x_lo + lx / world),
                            host_buffers=False)
        out = job.step_resident()
        res = {r: out[r].cpu().numpy() for r in RADII}
        otsu = job.otsu
        # the host-buffer path (one collective per radius, columns downloaded as soon as their radius is done): two
        # strips streamed from pinned memory must land in the host buffers bit-identical to the resident step
        job.host = own.cpu().pin_memory()
        job.out_host = {r: torch.empty(per, dtype=torch.float32).pin_memory() for r in RADII}
        job.steps_e2e(2)
        e2e_same = all(np.array_equal(job.out_host[r].numpy(), res[r], equal_nan=True) for r in RADII)
        ret[rank] = (res, job.checksum, otsu, e2e_same)
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_nccl_bit_identical_to_untiled():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from more3d_b200 import synthetic
    from more3d_b200.pipeline import multiscale_saliency
    from oracle import ref_path as rp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    torch.cuda.set_device(0)
    whole_pts = synthetic.hash_terrain_device(11, 0, TOTAL, TOTAL, noise=0.15)
    whole = multiscale_saliency(whole_pts, RADII, curvature_weight=0.5, keep=True, **PARAMS)
    per = TOTAL // world
    for r in RADII:
        sal = whole[r]["_saliency"].cpu().numpy()
        pc = whole[r]["_pcount"].cpu().numpy()
        assert np.isnan(sal).mean() < 1e-3 and np.unique(sal[~np.isnan(sal)]).size > 100        # not vacuous
        for k in range(world):
            assert np.array_equal(ret[k][0][r], sal[k * per:(k + 1) * per], equal_nan=True)    # bit-identical
            assert ret[k][1][r][0] == int(pc[k * per:(k + 1) * per].sum())
        assert ret[0][2][r] == ret[1][2][r] == rp.otsu_threshold(sal[~np.isnan(sal)])          # global Otsu
    assert ret[0][3] and ret[1][3]                                                             # host-buffer path
