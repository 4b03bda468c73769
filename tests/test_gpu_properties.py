"""GPU: size-independent properties at sizes the CPU oracle cannot finish (2 M points here; the same
checks run at the full BASELINE.json sizes in tools/bench_configs.py -> profiles/r01_configs_1_3_4.jsonl)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 2_000_000


@pytest.fixture(scope="module")
def cloud():
    from more3d_b200 import synthetic
    return synthetic.terrain(N, seed=17)


def test_neighbour_relation_is_symmetric_and_reflexive(cloud):
    import torch
    from more3d_b200.grid import UniformGrid, cell_for_radius
    d = torch.from_numpy(cloud).cuda()
    r = 0.75
    g = UniformGrid(d, cell_for_radius(r))
    off, idx = g.radius_csr(None, r)
    cnt = (off[1:] - off[:-1])
    assert int(cnt.min()) >= 1                                   # every point finds itself
    i = torch.arange(N, device="cuda", dtype=torch.int64)
    # j in N(i) <=> i in N(j):  sum_i sum_{j in N(i)} j == sum_j j * |N(j)|, and the same for squares
    assert int(idx.long().sum()) == int((i * cnt).sum())
    assert int((idx.long() ** 2).sum()) == int((i * i * cnt).sum())
    assert (int(cnt.sum()) - N) % 2 == 0                         # off-diagonal pairs come in twos
    # the same sets from the count kernel, the fused moments kernel and a grid with another cell size
    _, _, _, pc = g.normals_curvature(r, 0.0392, 4)
    assert torch.equal(pc.long(), cnt)
    g2 = UniformGrid(d, 1.37 * r)
    assert torch.equal(g2.radius_count(None, r).long(), cnt)
    # idempotence: same call, same bits
    off2, idx2 = g.radius_csr(None, r)
    assert torch.equal(off, off2) and torch.equal(idx, idx2)


def test_pipeline_outputs_are_reproducible_and_bounded(cloud):
    import torch
    from more3d_b200.pipeline import multiscale_saliency
    a = multiscale_saliency(cloud, (0.5, 1.0), keep=True)
    b = multiscale_saliency(cloud, (0.5, 1.0), keep=True)
    for r in (0.5, 1.0):
        for k in ("_nonparam_K", "_dk", "_dn", "_saliency", "_pcount"):
            assert torch.equal(a[r][k], b[r][k]), (r, k)        # bit-reproducible run to run
        nrm = a[r]["normals"]
        ok = ~torch.isnan(nrm[:, 0])
        assert float((nrm[ok].norm(dim=1) - 1).abs().max()) < 1e-12
        assert float((nrm[ok] * (-torch.from_numpy(cloud).cuda()[ok])).sum(1).min()) >= 0   # oriented to the origin
        K = a[r]["_nonparam_K"]
        nz = K[K != 0].abs()
        assert float(nz.min()) >= 0.0391 and float(a[r]["_dk"].min()) >= 0 and float(a[r]["_dn"].min()) >= 0
        mm = a[r]["_dkdn_minmax"].cpu().numpy()
        sal = a[r]["_saliency"]
        lo = 0.5 * (mm[0] + mm[2])                               # both normalised columns start at their min
        assert float(sal.min()) >= lo - 1e-6 and float(sal.max()) <= lo + 1.0 + 1e-6


def test_voxel_and_cells_conserve_the_cloud(cloud):
    import ctypes as C
    import torch
    from more3d_b200 import _lib
    from more3d_b200._lib import check, ptr, stream_ptr
    from more3d_b200.BallTreePointSet import BallTreePointSet
    lib = _lib.load()
    d = torch.from_numpy(cloud).cuda()
    out = torch.empty((N, 3), dtype=torch.float64, device="cuda")
    vop = torch.empty(N, dtype=torch.int32, device="cuda")
    nout = C.c_int64(0)
    check(lib.more3d_voxel_downsample(ptr(d), N, 2.0, ptr(out), None, ptr(vop), C.byref(nout), stream_ptr()))
    V = int(nout.value)
    cnt = torch.bincount(vop.long(), minlength=V).double()
    assert int(cnt.min()) >= 1 and int(cnt.sum()) == N
    total = (out[:V] * cnt[:, None]).sum(0)
    assert torch.allclose(total, d.sum(0), rtol=1e-12)           # sum_v count_v * mean_v == sum of points
    # each point lies inside its voxel's cube
    vmin = d.min(0).values - 1.0
    key = torch.floor((d - vmin) / 2.0)
    mean_key = torch.floor((out[:V] - vmin) / 2.0)
    assert torch.equal(mean_key[vop.long()], key)               # the mean stays in the (convex) voxel
    bt = BallTreePointSet(cloud, leaf_size=40)
    t = bt._hierarchy()
    for lod in (1.0, 5.0):
        cells = np.atleast_1d(bt.getSmallestNodesOfSize(lod))
        sizes = t.idx_end[cells] - t.idx_start[cells]
        assert sizes.sum() == N and np.all(np.diff(t.idx_start[cells]) == sizes[:-1])
        assert np.all((t.node_radius[cells] < lod) | (t.is_leaf[cells] == 1))
    assert np.array_equal(np.sort(t.idx_array), np.arange(N))    # idx_array is a permutation
