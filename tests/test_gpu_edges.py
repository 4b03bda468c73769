"""GPU: edge cases of the hot path -- duplicates (np.unique semantics), tiny / degenerate clouds,
empty and far-away queries, radius 0, ragged neighbourhoods below min_points, wide stencils."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_close(got, want, rtol=1e-4, atol=1e-9):
    return np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64)) <= atol + rtol * np.abs(np.asarray(want, np.float64))


def test_duplicate_points_follow_np_unique():
    """Coincident points: neighbour sets contain all copies, sizePoint() counts them, the dk/dn statistics
    run over the de-duplicated neighbourhood (DM_saliency.py:247, :283-284)."""
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature, DM_dk_dn
    from oracle import ref_path as rp
    base = synthetic.terrain(2500, seed=41, noise=0.15)
    rng = np.random.Generator(np.random.PCG64(7))
    dup = rng.choice(len(base), 300, replace=False)
    pts = np.vstack([base, base[dup], base[dup[:60]]])          # doubles and triples
    pts = pts[rng.permutation(len(pts))]
    r, rho = 0.6, 0.6 / np.sqrt(2)
    tree = rp.build_tree(pts)
    want = rp.query_radius(tree, pts, r)
    got = BallTreePointSet(pts).queryRadius(pts, r)
    for a, b in zip(got, want):
        assert np.array_equal(np.sort(a), np.sort(b))
    dm = PointCloudDM(pts)
    DM_nonparam_curvature(dm, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=r)
    nrm, K = dm.normals.cpu().numpy(), dm.get("_nonparam_K")
    K_or, cnt_or, _ = rp.curvature_kernel(pts, nrm, want, 0.05, 0.02, 4)
    assert np.array_equal(dm.get("_pcount"), cnt_or)
    assert (~rel_close(K, K_or)).sum() <= 2
    DM_dk_dn(dm, "sphere", rho, rho, sigma_normal=1e-3, sigma_curvature=0.01, radius=r)
    dk_or, dn_or = rp.saliency_kernel(pts, nrm, K, want, rho, 1e-3, 0.01, 0.05, 4)
    dk, dn = dm.get("_dk"), dm.get("_dn")
    assert (dk_or != 0).mean() > 0.05 and (dn_or != 0).mean() > 0.05          # the branches are exercised
    assert (~rel_close(dk, dk_or)).sum() <= 3 and (~rel_close(dn, dn_or)).sum() <= 3


def test_tiny_and_degenerate_clouds():
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature, DM_dk_dn, DM_saliency
    one = np.array([[1.0, 2.0, 3.0]])
    bt = BallTreePointSet(one)
    assert np.array_equal(bt.queryRadius(one, 0.5)[0], [0])
    assert np.array_equal(bt.queryRadius(np.array([9.0, 9.0, 9.0]), 0.5), np.zeros(0, np.int64))
    assert bt.query(one, 1)[0, 0] == 0
    # all points on one vertical line: a single xy column, huge z extent
    line = np.zeros((500, 3))
    line[:, 2] = np.arange(500) * 0.1
    bt = BallTreePointSet(line)
    got = bt.queryRadius(line, 0.25)
    assert [len(g) for g in got[:3]] == [3, 4, 5] and len(got[250]) == 5
    # isolated points: fewer than min_points neighbours -> K = dk = dn = 0, NaN normal below 3 neighbours
    sparse = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [10, 10, 0], [5, 5, 0.1]], dtype=np.float64)
    dm = PointCloudDM(sparse)
    DM_nonparam_curvature(dm, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=1.0)
    assert np.array_equal(dm.get("_pcount"), [1, 1, 1, 1, 1])
    assert np.all(np.isnan(dm.normals.cpu().numpy())) and np.all(dm.get("_nonparam_K") == 0)
    DM_dk_dn(dm, "sphere", 0.7, 0.7, radius=1.0)
    assert np.all(dm.get("_dk") == 0) and np.all(dm.get("_dn") == 0)
    DM_saliency(dm, 0.5)
    assert np.all(np.isnan(dm.get("_saliency")))          # 0/0 in the reference's normalisation as well


def test_query_shapes_radii_and_far_queries():
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from oracle import ref_path as rp
    pts = synthetic.terrain(5000, seed=3)
    bt = BallTreePointSet(pts)
    tree = rp.build_tree(pts)
    assert len(bt.queryRadius(np.zeros((0, 3)), 1.0)) == 0                   # empty query set
    zero = bt.queryRadius(pts[:50], 0.0)                                      # r = 0: the point itself
    assert all(np.array_equal(z, [i]) for i, z in enumerate(zero))
    q = np.vstack([pts[:20] + [0.3, -0.2, 0.05], [[1e6, 1e6, 0.0]], [[0.0, 0.0, 1e5]], pts.mean(0)[None] + [0, 0, 3.0]])
    for r in (0.05, 0.7, 3.0, 40.0):                                          # 40 m: stencil of >100 columns
        want = rp.query_radius(tree, q, r)
        got = bt.queryRadius(q, r)
        for a, b in zip(got, want):
            assert np.array_equal(np.sort(a), np.sort(b))
    big = bt.queryRadius(pts[7], 1e4)                                         # everything
    assert len(big) == len(pts)
    # kNN with k = n and queries far outside
    small = pts[:40]
    bs = BallTreePointSet(small)
    ki, kd = bs.query(np.array([[1e3, 1e3, 1e3]]), 40, return_distnaces=True)
    wi, wd = rp.query_knn(rp.build_tree(small), np.array([[1e3, 1e3, 1e3]]), 40, return_distnaces=True)
    assert np.array_equal(ki, wi) and np.array_equal(kd, wd)


def test_large_coordinates_keep_exact_sets():
    """UTM-like coordinates (1e6 m offsets): the fp32 filter works on grid-local coordinates and the shell is
    re-checked in fp64, so sets stay bit-exact."""
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from oracle import ref_path as rp
    pts = synthetic.terrain(6000, seed=19) + [4.5e5, 5.3e6, 1500.0]
    tree = rp.build_tree(pts)
    bt = BallTreePointSet(pts)
    for r in (0.3, 1.1):
        want = rp.query_radius(tree, pts, r)
        got = bt.queryRadius(pts, r)
        for a, b in zip(got, want):
            assert np.array_equal(np.sort(a), np.sort(b))


def test_tall_columns_vertical_structures():
    """A cloud with vertical structures (columns of > 128 points) switches the grid to z-sorted columns with
    a per-column z bracket; results must be the same sets / values as the thin path's rule."""
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature, DM_dk_dn
    from more3d_b200.grid import UniformGrid, cell_for_radius
    from oracle import ref_path as rp
    import torch
    rng = np.random.Generator(np.random.PCG64(11))
    tower = rng.random((60000, 3)) * [3.0, 3.0, 60.0] + [0.0, 0.0, -1030.0]         # 3 x 3 m footprint, 60 m high
    wall = np.column_stack([rng.random(20000) * 12.0 - 6.0, np.full(20000, 4.0) + rng.normal(0, 0.01, 20000),
                            rng.random(20000) * 25.0 - 1010.0])
    ground = synthetic.terrain(8000, seed=5, noise=0.05)
    pts = np.vstack([tower, wall, ground, tower[:500]])                              # + coincident copies
    pts = pts[rng.permutation(len(pts))]
    r = 0.5
    g = UniformGrid(torch.from_numpy(pts).cuda(), cell_for_radius(r))
    cnt = g.radius_count(None, r).cpu().numpy()
    tree = rp.build_tree(pts)
    sub = rng.choice(len(pts), 4000, replace=False)
    want = rp.query_radius(tree, pts[sub], r)
    assert np.array_equal(cnt[sub], [len(w) for w in want])
    bt = BallTreePointSet(pts)
    got = bt.queryRadius(pts[sub], r)
    for a, b in zip(got, want):
        assert np.array_equal(np.sort(a), np.sort(b))
    dm = PointCloudDM(pts)
    DM_nonparam_curvature(dm, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=r)
    rho = r / np.sqrt(2)
    DM_dk_dn(dm, "sphere", rho, rho, sigma_normal=1e-3, sigma_curvature=0.01, radius=r)
    nrm, K, pc = dm.normals.cpu().numpy(), dm.get("_nonparam_K"), dm.get("_pcount")
    dk, dn = dm.get("_dk"), dm.get("_dn")
    assert np.array_equal(pc, cnt)
    eps = rp.curvature_epsilon(0.05, 0.02)
    table = rp.chi2_table(0.05, 4096)
    full = rp.query_radius(tree, pts[sub[:600]], r)
    bad = 0
    for i, b in zip(sub[:600], full):
        k_or, _, _ = rp.curvature_point(pts[i], nrm[i].copy(), pts[b], eps, 4)
        a_dk, a_dn = rp.saliency_point(pts[i], nrm[i], float(K[i]), pts[b], nrm[b], K[b].astype(np.float64), rho, 1e-3,
                                       0.01, table, 4)
        ok = rel_close(K[i], np.float32(k_or)) and rel_close(dk[i], np.float32(a_dk)) and rel_close(dn[i], np.float32(a_dn))
        bad += 0 if ok else 1
    assert bad <= 3, bad


def test_external_queries_in_any_order():
    """External query sets are walked in grid-column order inside the library (>= 4096 queries); results must
    land at the caller's query index whatever order the queries arrive in."""
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from oracle import ref_path as rp
    rng = np.random.Generator(np.random.PCG64(23))
    pts = synthetic.terrain(30000, seed=29)
    q = pts[rng.permutation(len(pts))[:9000]] + rng.normal(0, 0.2, (9000, 3))
    q[::1000] += [500.0, -300.0, 0.0]                                          # some far outside the grid
    tree = rp.build_tree(pts)
    bt = BallTreePointSet(pts)
    want = rp.query_radius(tree, q, 0.8)
    got = bt.queryRadius(q, 0.8)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(np.sort(a), np.sort(b))
    ki, kd = bt.query(q, 8, return_distnaces=True)
    wi, wd = rp.query_knn(tree, q, 8, return_distnaces=True)
    assert np.array_equal(kd, wd)
    same = ki == wi
    assert np.all(same | (kd == np.roll(kd, 1, axis=1)) | (kd == np.roll(kd, -1, axis=1)))   # ties may swap
