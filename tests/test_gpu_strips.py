"""GPU: tile-and-halo on one GPU with G logical ranks == the untiled result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RADII = (0.5, 1.0)
PARAMS = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)


def test_band_select_is_stable_and_exact():
    import torch
    from more3d_b200.distributed import band_select, gather_rows
    rng = np.random.Generator(np.random.PCG64(4))
    pts = rng.random((100003, 3)) * [50.0, 20.0, 3.0]
    d = torch.from_numpy(pts).cuda()
    for lo, hi in ((-np.inf, 7.5), (42.25, np.inf), (10.0, 10.5), (60.0, 70.0)):
        idx = band_select(d, 0, lo, hi).cpu().numpy()
        want = np.nonzero((pts[:, 0] >= lo) & (pts[:, 0] < hi))[0]
        assert np.array_equal(idx, want)
        got = gather_rows(d, torch.from_numpy(want).cuda()).cpu().numpy()
        assert np.array_equal(got, pts[want])


@pytest.mark.parametrize("G", [2, 4])
def test_logical_strips_match_untiled(G):
    import torch
    from more3d_b200 import synthetic
    from more3d_b200.grid import CELL_MARGIN
    from more3d_b200.distributed import band_select, gather_rows, strip_saliency, strip_bounds
    from more3d_b200.pipeline import multiscale_saliency
    pts = synthetic.terrain(120000, seed=8)
    whole = multiscale_saliency(pts, RADII, curvature_weight=0.5, keep=True, **PARAMS)
    d = torch.from_numpy(pts).cuda()
    edges = strip_bounds(pts[:, 0].min(), pts[:, 0].max(), G)
    edges[0], edges[-1] = -np.inf, np.inf
    halo = 2.0 * max(RADII) * CELL_MARGIN
    owned, ghosts = [], []
    for k in range(G):
        oi = band_select(d, 0, edges[k], edges[k + 1])
        gi = torch.cat([band_select(d, 0, edges[k] - halo, edges[k]), band_select(d, 0, edges[k + 1], edges[k + 1] + halo)])
        owned.append(oi)
        ghosts.append(gi)
    # pass 1: local extrema; pass 2: blend with the global extrema (what the all-reduce delivers)
    local_mm = {r: [] for r in RADII}
    for k in range(G):
        it = iter(RADII)
        strip_saliency(gather_rows(d, owned[k]), gather_rows(d, ghosts[k]), RADII, PARAMS, 0.5,
                       reduce_minmax=lambda mm: local_mm[next(it)].append(mm.clone()))
    glob = {}
    for r in RADII:
        st = torch.stack(local_mm[r])
        glob[r] = torch.stack([st[:, 0].min(), st[:, 1].max(), st[:, 2].min(), st[:, 3].max()])
    for r in RADII:
        assert torch.equal(glob[r].cpu(), whole[r]["_dkdn_minmax"].cpu())
    for k in range(G):
        it = iter(RADII)
        out, pc = strip_saliency(gather_rows(d, owned[k]), gather_rows(d, ghosts[k]), RADII, PARAMS, 0.5,
                                 reduce_minmax=lambda mm: mm.copy_(glob[next(it)]))
        oi = owned[k].cpu().numpy()
        for r in RADII:
            assert np.array_equal(pc[r].tensor().cpu().numpy(), whole[r]["_pcount"].cpu().numpy()[oi])   # integer: exact
            a = out[r].cpu().numpy().astype(np.float64)
            b = whole[r]["_saliency"].cpu().numpy()[oi].astype(np.float64)
            bad = np.abs(a - b) > 1e-5 * np.abs(b) + 1e-7
            assert bad.sum() <= 2, (k, r, bad.sum())       # threshold straddlers under a different sum order


def test_device_otsu_matches_oracle():
    """DeviceOtsu (extrema, linspace edges and histogram on the device, all-reduces, one final wait) == the oracle's
    skimage.threshold_otsu restatement on the same column; a one-rank NCCL group stands in for the strips."""
    import torch
    import torch.distributed as dist
    from more3d_b200.distributed import DeviceOtsu
    from oracle import ref_path as rp
    created = False
    if not dist.is_initialized():
        import socket
        with socket.socket() as sock:                       # a free port, not a fixed one
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        try:
            dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1,
                                    device_id=torch.device("cuda", 0))
        except Exception as e:                              # no NCCL rendezvous on this box: not a parity failure
            pytest.skip("cannot create a one-rank NCCL group: %r" % (e,))
        created = True
    try:
        rng = np.random.default_rng(5)
        for col in (rng.random(200000).astype(np.float32) ** 4, (rng.normal(size=50001) * 3 - 1).astype(np.float32),
                    np.full(1000, 0.25, dtype=np.float32)):
            got = DeviceOtsu(torch.from_numpy(col).cuda()).result()
            assert got == rp.otsu_threshold(col)
        cols = [rng.random(30000).astype(np.float32) ** k for k in (1, 3, 6)]      # several columns in one job
        got = DeviceOtsu([torch.from_numpy(c).cuda() for c in cols]).results()
        assert got == [rp.otsu_threshold(c) for c in cols]
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("G", [2, 3])
def test_voxel_grid_across_logical_ranks_equals_single_gpu(G):
    """voxel_downsample_strips with G logical ranks (host threads) on an arbitrary, non voxel-aligned partition: the
    union of the ranks' voxels is the single-GPU result -- same voxel indices, bit-identical fp64 means."""
    import threading
    import torch
    from more3d_b200 import synthetic
    from more3d_b200.Downsample import voxel_downsample
    from more3d_b200.distributed import voxel_downsample_strips
    from more3d_b200.distributed_levelset import ThreadComm
    pts = synthetic.terrain(50000, seed=12)
    rng = np.random.default_rng(3)
    part = rng.integers(0, G, len(pts))                      # points scattered over the ranks at random
    part[pts[:, 0] < np.quantile(pts[:, 0], 0.3)] = 0        # ... plus a spatially coherent block
    for vs in (0.7, 3.0):
        want, wkeys = voxel_downsample(pts, vs, return_keys=True)
        hub = ThreadComm.Hub(G)
        out, err = [None] * G, []

        def worker(r):
            try:
                torch.cuda.set_device(0)
                idx = np.nonzero(part == r)[0]
                m, k = voxel_downsample_strips(ThreadComm(hub, r), pts[idx], vs, gid_own=idx)
                out[r] = (m.cpu().numpy(), k.cpu().numpy())
            except Exception as e:      # pragma: no cover
                err.append((r, repr(e)))
                hub.barrier.abort()

        th = [threading.Thread(target=worker, args=(r,)) for r in range(G)]
        [t.start() for t in th]
        [t.join(timeout=300) for t in th]
        assert not err, err
        means = np.concatenate([o[0] for o in out])
        keys = np.concatenate([o[1] for o in out])
        assert len(keys) == len(wkeys) and all(len(o[1]) > 0 for o in out)
        order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
        assert np.array_equal(keys[order], wkeys)             # voxel assignment: identical
        assert np.array_equal(means[order], want)             # members accumulated in the same order: identical means
        owners = [set(map(int, o[1][:, 0])) for o in out]     # voxel-aligned: x-index sets of the ranks are disjoint
        assert all(owners[a].isdisjoint(owners[b]) for a in range(G) for b in range(a + 1, G))
