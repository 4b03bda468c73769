"""GPU parity tests at the sizes BASELINE.json names (configs[0]-[3]); the small-size tests next to this file
prove parity against the verbatim goldens, these prove that nothing changes at scale: u32 positions, multi-million
column tables, 19-21-level trees, 8 GB operators.

The CPU side is the reference's own calls where they can run at size (sklearn BallTree build / query_radius /
query -- BallTreePointSet.py:35, :311, :345 -- and, through oracle/_ref, the UNMODIFIED Curvature_Kernel /
Saliency_Kernel callbacks on a sample of query points against the full-N tree), else the oracle restatement.
Tolerances as everywhere: sets / counts / selections / keys bit-exact, normals <= 1e-5 rad, K / dk / dn <= 1e-4
relative with threshold straddlers counted.
"""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.fullsize]

# MORE3D_FULLSIZE_SCALE=0.02 runs the same tests on 2 % of the points (a quick check of the test logic itself)
SCALE = float(os.environ.get("MORE3D_FULLSIZE_SCALE", "1"))


def sized(n):
    return max(int(n * SCALE), 20_000)

PARAMS = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
RTOL = 1e-4


def rel_bad(got, want, rtol=RTOL, atol=1e-9):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want) > atol + rtol * np.abs(want)


def angle(a, b):
    c = np.abs(np.sum(a * b, axis=1)) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    return np.arccos(np.clip(c, -1, 1))


def test_hash_terrain_device_matches_numpy_twin():
    """the counter-based input generator (bench / these tests): device == numpy twin, any id range"""
    from more3d_b200 import synthetic
    total = 3_000_000
    for start, count in ((0, 200_000), (1_234_567, 150_001), (total - 77, 77)):
        d = synthetic.hash_terrain_device(9, start, count, total).cpu().numpy()
        h = synthetic.hash_terrain_host(9, start, count, total)
        assert np.array_equal(d[:, :2], h[:, :2])                    # integer hashing + exact arithmetic
        assert np.abs(d[:, 2] - h[:, 2]).max() < 1e-9                # sin / cos / exp: libm vs CUDA, a few ulp
    # an id range is an x-strip
    lx, _ = synthetic.hash_tile_extent(total)
    d = synthetic.hash_terrain_device(9, total // 4, total // 4, total).cpu().numpy()
    assert d[:, 0].min() >= -0.25 * lx - 1e-9 and d[:, 0].max() <= 1e-9


def test_config0_query_radius_1m_all_rows_vs_sklearn():
    """configs[0]: 1 M points, r = 0.5 -- every one of the 1 M neighbour sets against sklearn's (the call the
    reference makes at BallTreePointSet.py:345)."""
    import torch
    from sklearn.neighbors import BallTree
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    n, r = sized(1_000_000), 0.5
    pts = synthetic.terrain(n, seed=1)
    bt = BallTreePointSet(pts, leaf_size=40)
    off, idx = bt.queryRadiusCSR(None, r)
    off, idx = off.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    tree = BallTree(pts, leaf_size=40)
    want = tree.query_radius(pts, r)
    cnt = np.fromiter((len(w) for w in want), np.int64, n)
    assert np.array_equal(np.diff(off), cnt)
    flat = np.concatenate(want)
    seg = np.repeat(np.arange(n), cnt)
    # canonical order inside every row, all rows at once
    assert np.array_equal(idx[np.lexsort((idx, seg))], flat[np.lexsort((flat, seg))])
    # the reference-shaped API on a slice of external queries (object array of int64 arrays)
    q = pts[::1000] + 1e-3
    got = bt.queryRadius(q, r)
    for a, b in zip(got, tree.query_radius(q, r)):
        assert a.dtype == np.int64 and np.array_equal(np.sort(a), np.sort(b))
    del bt
    torch.cuda.empty_cache()


def test_config1_saliency_10m_sample_vs_verbatim_reference():
    """configs[1]: 10 M points x 3 radii on the GPU; 6 000 sampled query points per radius against the full-N
    sklearn tree through the UNMODIFIED reference callbacks (oracle/_ref) -- or the oracle port if that is absent."""
    import torch
    from sklearn.neighbors import BallTree
    from more3d_b200 import synthetic
    from more3d_b200.pipeline import multiscale_saliency
    from oracle import ref_path as rp, ref_verbatim as rv
    n, radii, q = sized(10_000_000), (0.5, 0.75, 1.0), min(6000, sized(10_000_000) // 4)
    d = synthetic.hash_terrain_device(2, 0, n, n)
    res = multiscale_saliency(d, radii, curvature_weight=0.5, keep=True, **PARAMS)
    pts = d.cpu().numpy()
    tree = BallTree(pts, leaf_size=40)
    rng = np.random.Generator(np.random.PCG64(7))
    verbatim = rv.available()
    for r in radii:
        sub = np.sort(rng.choice(n, q, replace=False))
        nb = tree.query_radius(pts[sub], r)
        cols = {k: res[r][k] for k in ("_pcount", "_nonparam_K", "_dk", "_dn", "_saliency", "normals")}
        pc = cols["_pcount"].cpu().numpy()
        assert np.array_equal(pc[sub], np.fromiter((len(j) for j in nb), np.int64, q))      # counts: exact
        nrm = cols["normals"].cpu().numpy()
        K = cols["_nonparam_K"].cpu().numpy()
        n_or = np.empty((q, 3))
        for t, (i, j) in enumerate(zip(sub, nb)):
            n_or[t], _ = rp.pca_point(pts[i], pts[j])
        assert angle(nrm[sub], n_or).max() <= 1e-5
        if verbatim:
            K_ref, pc_ref, n_fl = rv.run_curvature(pts, n_or, nb, PARAMS["alpha"], PARAMS["roughness"],
                                                   PARAMS["min_points"], queries=sub)
        else:
            eps = rp.curvature_epsilon(PARAMS["alpha"], PARAMS["roughness"])
            out = [rp.curvature_point(pts[i], n_or[t], pts[j], eps, PARAMS["min_points"]) for t, (i, j) in enumerate(zip(sub, nb))]
            K_ref = np.array([o[0] for o in out], np.float32)
            n_fl = np.array([o[2] for o in out])
        live = pc[sub] >= PARAMS["min_points"]
        assert np.all(np.sum(nrm[sub] * n_fl, axis=1)[live] > 0)                            # same orientation
        bad = rel_bad(K[sub], K_ref)
        assert bad.sum() <= 4, (r, int(bad.sum()))                                          # eps straddlers only
        rho = r / np.sqrt(2.0)
        if verbatim:
            dk_ref, dn_ref = rv.run_dk_dn(pts, nrm, K, nb, rho, rho, PARAMS["sigma_normal"], PARAMS["sigma_curvature"],
                                          PARAMS["alpha"], PARAMS["min_points"], queries=sub)
        else:
            table = rp.chi2_table(PARAMS["alpha"], 4096)
            o = [rp.saliency_point(pts[i], nrm[i], K[i], pts[j], nrm[j], K[j], rho, PARAMS["sigma_normal"],
                                   PARAMS["sigma_curvature"], table, PARAMS["min_points"]) for i, j in zip(sub, nb)]
            dk_ref, dn_ref = np.array([a for a, _ in o]), np.array([b for _, b in o])
        dk, dn = cols["_dk"].cpu().numpy(), cols["_dn"].cpu().numpy()
        bad = rel_bad(dk[sub], dk_ref) | rel_bad(dn[sub], dn_ref)
        assert bad.sum() <= 4, (r, int(bad.sum()))
        assert (dk[sub] != 0).mean() > 0.02                                                 # not vacuous
        sal_or, _ = rp.saliency_blend(dk, dn, 0.5)                                          # whole columns (global min/max)
        assert np.array_equal(cols["_saliency"].cpu().numpy(), sal_or, equal_nan=True)
        assert np.isnan(sal_or).mean() < 1e-3                                               # not vacuous
    del res, d
    torch.cuda.empty_cache()


def _canon(idx_array, starts, n):
    """idx_array with every leaf segment sorted: equal iff all leaf memberships are equal"""
    seg = np.searchsorted(starts, np.arange(n), side="right")
    return idx_array[np.lexsort((idx_array, seg))]


def test_config1_two_pass_dkdn_equals_one_pass_10m():
    """configs[1] at full size: the default dk/dn (screening walk + refinement) against the one-pass
    kernel on all 3 x 10 M points -- identical neighbour counts and zero / NaN patterns (the integer decisions), values
    equal to the rounding of fp64 sums added in another order (observed: bit-identical float32 columns)."""
    import torch
    from more3d_b200 import synthetic, _lib
    from more3d_b200.pipeline import multiscale_saliency
    n = sized(10_000_000)
    d = synthetic.hash_terrain_device(2, 0, n, n)
    cols = {}
    try:
        for key, variant in (("one", 2), ("two", 0)):
            _lib.set_dkdn_variant(variant)
            res = multiscale_saliency(d, (0.5, 0.75, 1.0), curvature_weight=0.5, keep=True, **PARAMS)
            cols[key] = {r: {k: res[r][k].cpu().numpy() for k in ("_dk", "_dn", "_pcount", "_saliency")} for r in res}
            del res
    finally:
        _lib.set_dkdn_variant(0)
    for r in (0.5, 0.75, 1.0):
        a, b = cols["one"][r], cols["two"][r]
        assert np.array_equal(a["_pcount"], b["_pcount"])
        for k in ("_dk", "_dn"):
            assert np.array_equal(a[k] == 0, b[k] == 0) and np.array_equal(np.isnan(a[k]), np.isnan(b[k])), (r, k)
            nz = (a[k] != 0) & ~np.isnan(a[k])
            assert np.allclose(a[k][nz], b[k][nz], rtol=1e-6, atol=0), (r, k)
        fin = np.isfinite(a["_saliency"])
        assert np.array_equal(fin, np.isfinite(b["_saliency"]))
        assert np.allclose(a["_saliency"][fin], b["_saliency"][fin], rtol=1e-5, atol=1e-7)
        assert 0.02 < (b["_dk"] != 0).mean() < 0.5          # the refinement really ran on a share of the points


def test_config2_hierarchy_and_selection_10m_vs_sklearn():
    """configs[2]: the LOD tree of a 10 M-point cloud against sklearn's BallTree (node ranges, leaf membership,
    radii), the cell lists of every LOD and one saliency-driven selection against the oracle loop."""
    import torch
    from sklearn.neighbors import BallTree
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from more3d_b200.Downsample import saliency_downsample
    from oracle import ref_path as rp
    n = sized(10_000_000)
    pts = synthetic.hash_terrain_device(3, 0, n, n).cpu().numpy()
    bt = BallTreePointSet(pts, leaf_size=40)
    tr = bt._hierarchy()
    sk = BallTree(pts, leaf_size=40)
    _, sk_idx, nodes, _ = sk.get_arrays()
    assert tr.n_nodes == len(nodes)
    assert np.array_equal(tr.idx_start, nodes["idx_start"]) and np.array_equal(tr.idx_end, nodes["idx_end"])
    assert np.array_equal(tr.is_leaf, nodes["is_leaf"])
    leaves = np.flatnonzero(nodes["is_leaf"] == 1)
    starts = np.sort(nodes["idx_start"][leaves])
    assert np.array_equal(_canon(tr.idx_array, starts, n), _canon(np.asarray(sk_idx), starts, n))   # every leaf set
    np.testing.assert_allclose(tr.node_radius, nodes["radius"], rtol=1e-11, atol=0)
    for lod in (1, 2, 5, 7, 10):
        assert np.array_equal(np.atleast_1d(bt.getSmallestNodesOfSize(lod)), rp.lod_cells(sk, lod))
    rng = np.random.Generator(np.random.PCG64(11))
    sal = rng.random(n) ** 4
    got = saliency_downsample(pts, sal, 5.0, btP=bt)
    want_xyz, want_flag, _, _ = rp.saliency_downsample(pts, sal, 5.0, tree=sk)
    assert got.shape == (len(want_xyz), 4)
    # rows cell by cell; inside a cell the kept points follow idx_array order, which is set-equal, not order-equal
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0], a[:, 3]))]
    assert np.array_equal(key(got), key(np.column_stack([want_xyz, want_flag])))
    del bt, tr
    torch.cuda.empty_cache()


def test_config2_voxel_grid_50m_sampled_voxels():
    """configs[2]: 50 M points, voxel 2 m -- voxel count and 300 sampled voxels (indices + fp64 means, members
    summed in point order) against the oracle's arithmetic."""
    import torch
    from more3d_b200 import synthetic
    from more3d_b200.Downsample import voxel_downsample
    n, voxel = sized(50_000_000), 2.0
    d = synthetic.hash_terrain_device(3, 0, n, n)
    out, keys = voxel_downsample(d, voxel, return_keys=True, return_device=True)
    pts = d.cpu().numpy()
    del d
    vmin = pts.min(axis=0) - voxel * 0.5                       # oracle/ref_path.py:voxel_keys
    k = np.floor((pts - vmin) / voxel).astype(np.int64)
    dims = k.max(axis=0) + 1
    lin = (k[:, 0] * dims[1] + k[:, 1]) * dims[2] + k[:, 2]
    del k
    uniq = np.unique(lin)
    assert out.shape[0] == len(uniq)
    gk = keys.cpu().numpy()
    glin = (gk[:, 0] * dims[1] + gk[:, 1]) * dims[2] + gk[:, 2]
    assert np.array_equal(glin, uniq)                          # same voxels, sorted by voxel index
    rng = np.random.Generator(np.random.PCG64(5))
    pick = rng.choice(len(uniq), min(300, len(uniq)), replace=False)
    members = np.flatnonzero(np.isin(lin, uniq[pick]))
    got = out.cpu().numpy()
    order = np.argsort(lin[members], kind="stable")            # point order inside every voxel
    ml, mp = lin[members][order], pts[members][order]
    bounds = np.flatnonzero(np.diff(ml)) + 1
    for seg_l, seg_p in zip(np.split(ml, bounds), np.split(mp, bounds)):
        s = np.zeros(3)
        for p in seg_p:                                        # sequential fp64 accumulation, like np.add.at
            s = s + p
        row = np.searchsorted(uniq, seg_l[0])
        assert np.array_equal(got[row], s / len(seg_p))
    del out, keys
    torch.cuda.empty_cache()


def test_config3_levelset_1m_10_iterations_vs_oracle():
    """configs[3] at 1 M points: kNN table against sklearn on a query sample, then 10 Chan-Vese iterations against
    the numpy restatement of levelsets_func.Solver.run started from the same frames and weights."""
    import torch
    from sklearn.neighbors import BallTree
    from more3d_b200 import synthetic
    from more3d_b200 import levelsets_func as ls
    from oracle import ref_levelset as rl
    n, h, k = sized(1_000_000), 0.5, 7
    pts = synthetic.hash_terrain_device(4, 0, n, n).cpu().numpy()
    pts -= pts.mean(axis=0)
    nb, nrm, (t1, t2) = ls.build_neighborhoods(pts.copy(), h, k)
    tree = BallTree(pts, leaf_size=64)                         # levelsets_func.py:158
    sub = np.arange(0, n, 23)
    want = tree.query(pts[sub], k=2 * k, return_distance=False, sort_results=True)
    assert np.array_equal(nb[sub], want[:, 1:k + 1])           # levelsets_func.py:163, :172
    rng = np.random.Generator(np.random.PCG64(3))
    zeta = np.abs(np.sin(pts[:, :1] / 3.0) * np.cos(pts[:, 1:2] / 2.0)) + 0.05 * rng.random((n, 1))
    # the oracle's weights (numpy's SIMD pow is host dependent) are handed to the CUDA solver, so both
    # sides iterate on identical operands; the frames are the CUDA ones
    cw, cD = rl.wls_operator(pts, nb, t1, t2, 3 * h)
    solver = ls.Solver(h, zeta.copy(), pts, nb, (t1, t2), "chan_vese", weights=cw)
    ls.normalize(solver.zeta)
    solver.initialize_from_voxels(2.0)
    w = cw
    phi0 = solver.phi.copy()
    act0 = rl.zero_crossings(phi0, nb, np.arange(n))
    assert np.array_equal(solver.active, act0)
    st = dict(points=pts, nbrs=nb, t1=t1, t2=t2, w=w, D=cD, h=h, zeta=solver.zeta.copy(), phi=phi0.copy(), active=act0.copy(),
              last_active=None, alias=True)
    st["last_active"] = st["active"]
    run = dict(stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
    solver.run(10, **run)
    rl.run(st, 10, **run)
    assert np.array_equal(solver.active, st["active"])         # active set: exact
    np.testing.assert_allclose(solver.phi, st["phi"], rtol=1e-8, atol=1e-10)
    del solver
    torch.cuda.empty_cache()
