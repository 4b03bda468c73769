"""GPU parity tests of the saliency path (grid -> radius query -> normals/curvature -> dk/dn ->
blend/histogram) through the C-ABI, against the verbatim-reference golden vectors and the oracle.

Tolerances (BASELINE.json north_star): index sets / counts bit-exact; normals <= 1e-5 rad;
curvature, dk, dn, saliency <= 1e-4 relative.  A mismatch is excused only when the point sits on a
decision threshold of the reference (|sum_proj| ~ eps, chi^2 ~ table, 60 % rule) -- those are counted
and must stay a tiny fraction.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SAL_CASES = ["terrain_r05", "terrain_r10", "rough_r05", "boulders_r075", "dups_sparse_r06"]
RTOL = 1e-4


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def nbr_lists(g):
    off, idx = g["nbr_off"], g["nbr_idx"]
    return [idx[off[i]:off[i + 1]].astype(np.int64) for i in range(len(off) - 1)]


def angle(a, b):
    c = np.abs(np.sum(a * b, axis=1)) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    return np.arccos(np.clip(c, -1, 1))


def rel_close(got, want, rtol=RTOL, atol=1e-9):
    return np.abs(got.astype(np.float64) - want.astype(np.float64)) <= atol + rtol * np.abs(want.astype(np.float64))


@pytest.mark.parametrize("name", SAL_CASES + ["lattice_r5"])
def test_query_radius_sets_bit_exact(golden_dir, name):
    from more3d_b200.BallTreePointSet import BallTreePointSet
    g = load(golden_dir, name)
    pts = g["points"]
    q = g["queries"] if "queries" in g.files else pts
    bt = BallTreePointSet(pts, leaf_size=40)
    got = bt.queryRadius(q, float(g["radius"]))
    want = nbr_lists(g)
    assert got.dtype == object and len(got) == len(want)
    for a, b in zip(got, want):
        assert a.dtype == np.int64
        assert np.array_equal(np.sort(a), b)
    if name == "lattice_r5":
        one = bt.queryRadius(pts[100], 5.0)          # 1-D input -> bare index array
        assert np.array_equal(np.sort(one), g["single"])
        one = bt.queryRadius([list(pts[100])], 5.0)   # list input
        assert np.array_equal(np.sort(one[0]), g["single"])


def test_query_radius_boundary_ulps():
    """r = nextafter(d, 0), d, nextafter(d, inf) for true pair distances d (inclusive <=, fp64 rule)."""
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from more3d_b200 import synthetic
    from oracle import ref_path as rp
    pts = synthetic.terrain(4000, seed=77)
    bt = BallTreePointSet(pts)
    rng = np.random.Generator(np.random.PCG64(3))
    for qi in rng.integers(0, len(pts), 12):
        d2 = ((pts[qi, 0] - pts[:, 0]) ** 2 + (pts[qi, 1] - pts[:, 1]) ** 2) + (pts[qi, 2] - pts[:, 2]) ** 2
        for target in np.sort(d2)[[5, 17, 40]]:
            d = np.sqrt(target)
            for r in (np.nextafter(d, 0), d, np.nextafter(d, np.inf)):
                want = rp.query_radius_bruteforce(pts, pts[qi], r)[0]
                got = bt.queryRadius(pts[qi], r)
                assert np.array_equal(np.sort(got), want)


def test_query_knn_matches_balltree(golden_dir):
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from more3d_b200 import synthetic
    from oracle import ref_path as rp
    g = load(golden_dir, "lattice_r5")
    bt = BallTreePointSet(g["points"])
    q = g["queries"][:50] + 0.25
    ki, kd = bt.query(q, 5, return_distnaces=True)
    assert np.array_equal(kd, g["knn_d"])
    # lattice ties: same distances; the nearest point (unique) must be the same index
    assert np.array_equal(ki[:, 0], g["knn_i"][:, 0])
    pts = synthetic.terrain(20000, seed=5)
    bt = BallTreePointSet(pts)
    tree = rp.build_tree(pts)
    for k in (1, 14, 40):
        want_i, want_d = rp.query_knn(tree, pts[:3000], k, return_distnaces=True)
        got_i, got_d = bt.query(pts[:3000], k, return_distnaces=True)
        assert np.array_equal(got_i, want_i)
        assert np.array_equal(got_d, want_d)
    far = np.array([[1e4, -1e4, 0.0], [0.0, 0.0, 5e3]])
    want_i = rp.query_knn(tree, far, 3)
    assert np.array_equal(bt.query(far, 3), want_i)


@pytest.mark.parametrize("name", SAL_CASES)
def test_normals_and_curvature(golden_dir, name):
    import torch
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature
    g = load(golden_dir, name)
    r = float(g["radius"])
    # (1) fused: normals computed at the same radius
    dm = PointCloudDM(g["points"])
    DM_nonparam_curvature(dm, "sphere", roughness=float(g["roughness"]), alpha=float(g["alpha"]),
                          min_points=int(g["min_points"]), radius=r)
    pc = dm.get("_pcount")
    assert np.array_equal(pc, g["pcount"])
    nrm = dm.normals.cpu().numpy()
    has = ~np.isnan(g["normals"][:, 0])
    assert np.array_equal(np.isnan(nrm[:, 0]), ~has)
    ang = angle(nrm[has], g["normals"][has])
    assert ang.max() <= 1e-5, ang.max()
    # orientation: where the reference flipped toward the origin, so did we (sign now fixed)
    live = has & (g["pcount"] >= int(g["min_points"]))
    assert np.all(np.sum(nrm[live] * g["normals"][live], axis=1) > 0)
    ratio = dm.get("_lam_ratio")
    assert np.all(np.abs(ratio[has] - g["lam_ratio"][has]) <= 1e-4 * g["lam_ratio"][has] + 1e-12)
    K = dm.get("_nonparam_K")
    assert K.dtype == np.float32
    ok = rel_close(K, g["K"])
    eps = 1.959963984540054 * float(g["roughness"])
    big = np.maximum(np.abs(K), np.abs(g["K"]))
    straddle = ~ok & (np.abs(big - eps) <= 1e-6 * eps)
    assert np.all(ok | straddle), (np.nonzero(~(ok | straddle))[0][:10], K[~ok][:10], g["K"][~ok][:10])
    assert straddle.sum() <= 2
    # (2) normals supplied as a precondition (the reference's own usage): bit-exact flips
    dm2 = PointCloudDM(g["points"], normals=g["normals_in"])
    DM_nonparam_curvature(dm2, "sphere", roughness=float(g["roughness"]), alpha=float(g["alpha"]),
                          min_points=int(g["min_points"]), radius=r)
    assert np.array_equal(dm2.normals.cpu().numpy(), g["normals"], equal_nan=True)
    K2 = dm2.get("_nonparam_K")
    ok2 = rel_close(K2, g["K"])
    assert (~ok2).sum() <= 2


@pytest.mark.parametrize("name", SAL_CASES)
def test_dk_dn_and_saliency(golden_dir, name):
    from more3d_b200.DM_saliency import PointCloudDM, DM_dk_dn, DM_saliency, DM_saliency_stats
    from oracle import ref_path as rp
    g = load(golden_dir, name)
    r = float(g["radius"])
    dm = PointCloudDM(g["points"], normals=g["normals"])
    import torch
    dm.attrs["_nonparam_K"] = torch.from_numpy(g["K"]).cuda()
    DM_dk_dn(dm, "sphere", float(g["rho"]), float(g["rho"]), sigma_normal=float(g["sigma_n"]),
             sigma_curvature=float(g["sigma_k"]), alpha=float(g["alpha"]), min_points=int(g["min_points"]),
             radius=r)
    dk, dn = dm.get("_dk"), dm.get("_dn")
    ok_k, ok_n = rel_close(dk, g["dk"]), rel_close(dn, g["dn"])
    # threshold straddlers: one side exactly 0, the other not
    bad_k = ~ok_k & ~((dk == 0) ^ (g["dk"] == 0))
    bad_n = ~ok_n & ~((dn == 0) ^ (g["dn"] == 0))
    assert bad_k.sum() == 0 and bad_n.sum() == 0, (dk[bad_k][:5], g["dk"][bad_k][:5], dn[bad_n][:5], g["dn"][bad_n][:5])
    assert (~ok_k).sum() <= 3 and (~ok_n).sum() <= 3, ((~ok_k).sum(), (~ok_n).sum())
    print(name, "nonzero dk %.3f dn %.3f" % ((dk != 0).mean(), (dn != 0).mean()))
    # DM_saliency on the GOLDEN dk/dn so the blend is compared bit-for-bit with the oracle
    dm.attrs["_dk"] = torch.from_numpy(g["dk"]).cuda()
    dm.attrs["_dn"] = torch.from_numpy(g["dn"]).cuda()
    dm._minmax = None
    DM_saliency(dm, curvature_weight=0.5)
    sal = dm.get("_saliency")
    want, mm = rp.saliency_blend(g["dk"], g["dn"], 0.5)
    assert np.array_equal(dm.get("_dkdn_minmax"), np.array(mm, np.float32))
    assert np.array_equal(sal, want, equal_nan=True)
    if np.all(np.isfinite(want)):
        st = DM_saliency_stats(dm)
        assert np.allclose(st, rp.saliency_stats(want), rtol=1e-6)
        assert np.array_equal(dm.get("_saliency_minmax"), np.array([want.min(), want.max()], np.float32))


@pytest.mark.parametrize("params", [dict(sigma_normal=0.01, sigma_curvature=0.01, noise=0.02, boulders=0),
                                    dict(sigma_normal=1e-3, sigma_curvature=0.01, noise=0.05, boulders=3),
                                    dict(sigma_normal=0.05, sigma_curvature=0.1, noise=0.1, boulders=2)])
def test_two_pass_dkdn_equals_one_pass(params):
    """The default dk/dn (screening walk + refinement of the listed points) against the one-pass kernel on the same
    cloud, with coincident points, rough relief (the refinement's whole-expression path) and three radii on one grid:
    identical zero / NaN patterns and neighbour counts (integer decisions), values equal up to the rounding of fp64 sums
    that are added in another order."""
    import torch
    from more3d_b200 import synthetic, _lib
    from more3d_b200.pipeline import multiscale_saliency
    rng = np.random.Generator(np.random.PCG64(31))
    pts = synthetic.terrain(150000, seed=41, noise=params["noise"], boulders=params["boulders"])
    pts = np.vstack([pts, pts[rng.choice(len(pts), 3000, replace=False)]])           # coincident copies
    pts = pts[rng.permutation(len(pts))]
    d = torch.from_numpy(pts).cuda()
    kw = dict(roughness=0.02, alpha=0.05, sigma_normal=params["sigma_normal"], sigma_curvature=params["sigma_curvature"],
              min_points=4, keep=True)
    out = {}
    try:
        # 2: one pass; 0 with the fallback gate wide open (every listed point goes through the refinement kernels, also
        # where most of the cloud is listed) and with the default gate (rough clouds fall back to the one-pass kernel)
        for key, variant, gate in (("one", 2, 0.45), ("two", 0, 100.0), ("two_gated", 0, 0.45)):
            _lib.set_dkdn_variant(variant)
            _lib.set_dkdn_gate(gate)
            res = multiscale_saliency(d, (0.5, 0.75, 1.0), **kw)
            out[key] = {r: {k: v.cpu().numpy() for k, v in res[r].items() if k in ("_dk", "_dn", "_pcount", "_saliency")}
                        for r in res}
    finally:
        _lib.set_dkdn_variant(0)
        _lib.set_dkdn_gate(0.45)
    for r, key in [(r, key) for r in (0.5, 0.75, 1.0) for key in ("two", "two_gated")]:
        a, b = out["one"][r], out[key][r]
        assert np.array_equal(a["_pcount"], b["_pcount"])
        for k in ("_dk", "_dn"):
            assert np.array_equal(a[k] == 0, b[k] == 0), (r, k)
            assert np.array_equal(np.isnan(a[k]), np.isnan(b[k])), (r, k)
            nz = (a[k] != 0) & ~np.isnan(a[k])
            assert np.allclose(a[k][nz], b[k][nz], rtol=1e-6, atol=0), (r, k)
        fin = np.isfinite(a["_saliency"])
        assert np.array_equal(fin, np.isfinite(b["_saliency"]))
        assert np.allclose(a["_saliency"][fin], b["_saliency"][fin], rtol=1e-5, atol=1e-7)
        print(r, params, "nonzero dk %.3f dn %.4f" % ((b["_dk"] != 0).mean(), (b["_dn"] != 0).mean()))


def test_histogram_matches_numpy():
    import torch
    import ctypes as C
    from more3d_b200 import _lib
    from more3d_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    rng = np.random.Generator(np.random.PCG64(9))
    v = np.concatenate([rng.random(100000) ** 3, np.linspace(0, 1, 257), [0.0, 1.0]]).astype(np.float32)
    counts, edges = np.histogram(v.astype(np.float64), bins=256)
    dv = torch.from_numpy(v).cuda()
    de = torch.from_numpy(edges).cuda()
    hist = torch.empty(256, dtype=torch.int64, device="cuda")
    check(lib.more3d_histogram256(ptr(dv), 4, len(v), ptr(de), ptr(hist), stream_ptr()))
    assert np.array_equal(hist.cpu().numpy(), counts)


def test_plane_known_answers():
    from more3d_b200 import synthetic
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature, DM_dk_dn
    pts = synthetic.plane(20000, 3, normal=(0.1, -0.2, 1.0), extent=30.0)
    dm = PointCloudDM(pts)
    DM_nonparam_curvature(dm, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=0.8)
    n0 = np.array([0.1, -0.2, 1.0]) / np.linalg.norm([0.1, -0.2, 1.0])
    nrm = dm.normals.cpu().numpy()
    ok = dm.get("_pcount") >= 4
    assert np.all(np.abs(np.abs(nrm[ok] @ n0) - 1) < 1e-10)
    assert np.all(dm.get("_lam_ratio")[ok] < 1e-10)
    assert np.all(dm.get("_nonparam_K") == 0)
    DM_dk_dn(dm, "sphere", 0.8 / np.sqrt(2), 0.8 / np.sqrt(2), radius=0.8)
    assert np.all(dm.get("_dk") == 0) and np.all(dm.get("_dn") == 0)


def test_mid_size_against_oracle_subsample():
    """200 k points: CUDA path vs the oracle evaluated on a random subsample of query points."""
    from more3d_b200 import synthetic
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature, DM_dk_dn
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from oracle import ref_path as rp
    pts = synthetic.terrain(200000, seed=2)
    r, rho = 0.75, 0.75 / np.sqrt(2)
    dm = PointCloudDM(pts)
    DM_nonparam_curvature(dm, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=r)
    DM_dk_dn(dm, "sphere", rho, rho, sigma_normal=0.01, sigma_curvature=0.01, radius=r)
    nrm, K = dm.normals.cpu().numpy(), dm.get("_nonparam_K")
    dk, dn, pc = dm.get("_dk"), dm.get("_dn"), dm.get("_pcount")
    tree = rp.build_tree(pts)
    rng = np.random.Generator(np.random.PCG64(1))
    sub = rng.choice(len(pts), 1500, replace=False)
    nb = rp.query_radius(tree, pts[sub], r)
    bt = BallTreePointSet(pts)
    got = bt.queryRadius(pts[sub], r)
    eps = rp.curvature_epsilon(0.05, 0.02)
    table = rp.chi2_table(0.05, 4096)
    nbad = 0
    for a, b, i in zip(got, nb, sub):
        assert np.array_equal(np.sort(a), np.sort(b))
        assert pc[i] == len(b)
        n_or, _ = rp.pca_point(pts[i], pts[b])
        assert angle(n_or[None], nrm[i][None])[0] <= 1e-5
        k_or, _, n_fl = rp.curvature_point(pts[i], nrm[i].copy(), pts[b], eps, 4)
        if not rel_close(np.float32(K[i]), np.float32(k_or)):
            assert abs(max(abs(K[i]), abs(k_or)) - eps) < 1e-6 * eps
            nbad += 1
        a_dk, a_dn = rp.saliency_point(pts[i], nrm[i], float(K[i]), pts[b], nrm[b], K[b].astype(np.float64),
                                       rho, 0.01, 0.01, table, 4)
        if not (rel_close(dk[i], np.float32(a_dk)) and rel_close(dn[i], np.float32(a_dn))):
            nbad += 1
    assert nbad <= 3, nbad


def test_gridorder_columns_identical_to_point_order():
    """The grid-order variants (coalesced writes + lazy un-permutation) return bit-identical columns, the
    un-permute C entry point handles 4 / 8 / 24-byte elements, and a column survives the grid's destruction."""
    import torch
    from more3d_b200 import synthetic
    from more3d_b200.grid import UniformGrid, cell_for_radius
    from more3d_b200.DM_saliency import _chi2_table, PointCloudDM, DM_nonparam_curvature, DM_dk_dn, DM_saliency
    from more3d_b200.grid import GridColumn
    pts = synthetic.terrain(20000, seed=11)
    pts = pts[np.random.default_rng(0).permutation(len(pts))]
    r, rho = 0.6, 0.6 / np.sqrt(2)
    d = torch.from_numpy(pts).cuda()
    g = UniformGrid(d, cell_for_radius(r))
    table = _chi2_table(0.05, 1024, d.device)
    n0, ratio0, K0, pc0 = g.normals_curvature(r, 0.0392, 4)
    dk0, dn0, mm0 = g.dk_dn(r, n0, K0, rho, 0.01, 0.01, table, 4)
    n1, ratio1, K1, pc1 = g.normals_curvature_gridorder(r, 0.0392, 4)
    dk1, dn1, mm1 = g.dk_dn_gridorder(r, rho, 0.01, 0.01, table, 4)
    order = g.order().cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(len(pts)))
    assert np.array_equal(n1.data.cpu().numpy(), n0.cpu().numpy()[order], equal_nan=True)   # row i <-> point order[i]
    g.close()                                                                                  # columns outlive the grid
    for a, b in ((n0, n1), (ratio0, ratio1), (K0, K1), (pc0, pc1), (dk0, dk1), (dn0, dn1)):
        assert isinstance(b, GridColumn) and b.clean()
        assert np.array_equal(a.cpu().numpy(), b.tensor().cpu().numpy(), equal_nan=True)
    assert np.array_equal(mm0.cpu().numpy(), mm1.cpu().numpy())
    # through the reference-named API: default (grid-order inside) == columns handed over in point order
    dm = PointCloudDM(pts)
    DM_nonparam_curvature(dm, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=r)
    assert isinstance(dm.attrs.raw("_nonparam_K"), GridColumn)
    DM_dk_dn(dm, "sphere", rho, rho, sigma_normal=0.01, sigma_curvature=0.01, radius=r)
    DM_saliency(dm, curvature_weight=0.5)
    assert isinstance(dm.attrs.raw("_saliency"), GridColumn)
    dm2 = PointCloudDM(pts, normals=dm.normals)
    dm2.attrs["_nonparam_K"] = dm.attrs["_nonparam_K"].clone()
    DM_dk_dn(dm2, "sphere", rho, rho, sigma_normal=0.01, sigma_curvature=0.01, radius=r)
    DM_saliency(dm2, curvature_weight=0.5)
    assert not isinstance(dm2.attrs.raw("_saliency"), GridColumn)
    for k in ("_dk", "_dn", "_saliency"):
        assert np.array_equal(dm.get(k), dm2.get(k), equal_nan=True), k
    # an in-place edit of a materialised column invalidates the cached feature records: dk/dn must see it
    dm.attrs["_nonparam_K"].mul_(2.0)
    DM_dk_dn(dm, "sphere", rho, rho, sigma_normal=0.01, sigma_curvature=0.01, radius=r)
    dm2.attrs["_nonparam_K"].mul_(2.0)
    DM_dk_dn(dm2, "sphere", rho, rho, sigma_normal=0.01, sigma_curvature=0.01, radius=r)
    assert np.array_equal(dm.get("_dk"), dm2.get("_dk"), equal_nan=True)
    assert not np.array_equal(dm.get("_dk"), dk0.cpu().numpy())


def test_dm_functions_take_a_las_path(tmp_path):
    """DM_* accept the path of a LAS file where the reference takes an ODM path (DM_saliency.py:326)."""
    from more3d_b200 import synthetic, io_formats
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature
    pts = synthetic.terrain(3000, seed=5)
    path = str(tmp_path / "tile.las")
    io_formats.write_las(path, pts, scale=(1e-4, 1e-4, 1e-4))
    dm = DM_nonparam_curvature(path, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=0.5)
    q = io_formats.parse_las(path)
    ref = PointCloudDM(np.column_stack((q.x, q.y, q.z)))
    DM_nonparam_curvature(ref, "sphere", roughness=0.02, alpha=0.05, min_points=4, radius=0.5)
    assert np.array_equal(dm.get("_pcount"), ref.get("_pcount"))
    assert np.array_equal(dm.get("_nonparam_K"), ref.get("_nonparam_K"))


def test_tile_stream_matches_single_calls():
    """multiscale_saliency_tiles (uploads / downloads overlapped with compute) returns, per tile, exactly what a
    plain multiscale_saliency call on that tile returns, and fills the pinned output buffers."""
    import torch
    from more3d_b200 import synthetic
    from more3d_b200.pipeline import multiscale_saliency, multiscale_saliency_tiles
    radii = (0.5, 0.75)
    tiles = [torch.from_numpy(synthetic.terrain(20000 + 3000 * i, seed=20 + i, noise=0.06)).pin_memory()
             for i in range(3)]
    outs = [{r: torch.empty(t.shape[0], dtype=torch.float32).pin_memory() for r in radii} for t in tiles]
    kw = dict(curvature_weight=0.5, roughness=0.02, sigma_normal=1e-3, sigma_curvature=0.01)   # non-trivial dn, dk
    finite = 0
    res = multiscale_saliency_tiles(tiles, radii, out_hosts=outs, **kw)
    assert len(res) == 3
    for t, o, got in zip(tiles, outs, res):
        want = multiscale_saliency(t.numpy(), radii, **kw)
        for r in radii:
            w = want[r].cpu().numpy()
            finite += int(np.isfinite(w).all() and (w != w[0]).any())
            assert np.array_equal(got[r].cpu().numpy(), w, equal_nan=True)
            assert np.array_equal(o[r].numpy(), w, equal_nan=True)
    assert finite > 0        # (an all-zero dn or dk column normalises to NaN, DM_saliency.py:577-578)
    assert multiscale_saliency_tiles([], radii) == []
    plain = multiscale_saliency_tiles([tiles[0].numpy()], radii, **kw)           # pageable numpy input works too
    assert np.array_equal(plain[0][0.5].cpu().numpy(), res[0][0.5].cpu().numpy(), equal_nan=True)


def test_attribute_container_round_trip(tmp_path, golden_dir):
    """f-4: the on-disk attribute container (PointCloudDM.save_as / load) carries every column of the reference's ODM
    layouts (DM_saliency.py:377-384, :486-494) -- NormalX/Y/Z, _nonparam_K, _pcount, _dk, _dn, the normalised pair,
    _saliency and Classification -- with their storage types, and the pipeline resumes from a reloaded file."""
    import torch
    from more3d_b200.DM_saliency import PointCloudDM, DM_nonparam_curvature, DM_dk_dn, DM_saliency, DM_saliency_stats
    g = load(golden_dir, "rough_r05")
    r, rho = float(g["radius"]), float(g["rho"])
    dm = PointCloudDM(g["points"])
    cls = (np.arange(len(g["points"])) % 7).astype(np.uint8)
    dm.attrs["Classification"] = torch.from_numpy(cls).cuda()
    DM_nonparam_curvature(dm, "sphere", roughness=float(g["roughness"]), alpha=float(g["alpha"]),
                          min_points=int(g["min_points"]), radius=r)
    path1 = str(tmp_path / "stage1.npz")
    dm.save_as(path1)
    # second stage from the FILE, as the reference chains its stages through the ODM (DM_saliency.py:362, :415, :471)
    dm2 = DM_dk_dn(path1, "sphere", rho, rho, sigma_normal=float(g["sigma_n"]), sigma_curvature=float(g["sigma_k"]),
                   alpha=float(g["alpha"]), min_points=int(g["min_points"]), radius=r)
    DM_saliency(dm2, curvature_weight=0.5)
    path2 = str(tmp_path / "stage2.npz")
    dm2.save_as(path2)
    dm3 = PointCloudDM.load(path2)
    assert np.array_equal(dm3.xyz.cpu().numpy(), g["points"])
    want_types = {"_nonparam_K": np.float32, "_pcount": np.int32, "_dk": np.float32, "_dn": np.float32,
                  "_saliency": np.float32, "Classification": np.uint8}
    for name, t in want_types.items():
        a = dm3.get(name)
        assert a.dtype == t, (name, a.dtype)
        assert np.array_equal(a, dm2.get(name), equal_nan=True)
    for ax in "XYZ":
        assert np.array_equal(dm3.get("Normal" + ax), dm2.get("Normal" + ax), equal_nan=True)
    assert np.array_equal(dm3.get("Classification"), cls)
    # values are those of the in-memory pipeline / the verbatim golden
    assert np.array_equal(dm3.get("_pcount"), g["pcount"])
    bad = ~rel_close(dm3.get("_dk"), g["dk"]) | ~rel_close(dm3.get("_dn"), g["dn"])
    assert bad.sum() <= 3
    st = DM_saliency_stats(path2)
    s = dm3.get("_saliency").astype(np.float64)
    assert st[0] == s.min() and st[1] == s.max()
