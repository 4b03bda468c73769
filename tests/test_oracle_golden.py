"""CPU: the oracle restatement (oracle/ref_path.py) against golden vectors produced by executing
the reference verbatim (tools/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import ref_path as rp

SAL_CASES = ["terrain_r05", "terrain_r10", "rough_r05", "boulders_r075", "dups_sparse_r06"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def nbr_lists(g):
    off, idx = g["nbr_off"], g["nbr_idx"]
    return [idx[off[i]:off[i + 1]].astype(np.int64) for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", SAL_CASES)
def test_radius_sets_match_reference(golden_dir, name):
    g = load(golden_dir, name)
    pts, r = g["points"], float(g["radius"])
    tree = rp.build_tree(pts)
    got = rp.query_radius(tree, pts, r)
    want = nbr_lists(g)
    for a, b in zip(got, want):
        assert np.array_equal(np.sort(a), b)
    # the pure accept rule == the ball tree (bulk-accept / prune shortcuts never change a set)
    bf = rp.query_radius_bruteforce(pts, pts[:200], r)
    for a, b in zip(bf, want[:200]):
        assert np.array_equal(np.sort(a), b)


@pytest.mark.parametrize("name", SAL_CASES)
def test_curvature_kernel_bit_exact(golden_dir, name):
    g = load(golden_dir, name)
    K, cnt, nrm = rp.curvature_kernel(g["points"], g["normals_in"], nbr_lists(g), float(g["alpha"]),
                                      float(g["roughness"]), int(g["min_points"]))
    assert np.array_equal(cnt, g["pcount"])
    assert np.array_equal(nrm, g["normals"], equal_nan=True)
    # same expressions up to summation association -> a few ulp of f32
    np.testing.assert_allclose(K, g["K"], rtol=1e-6, atol=0)
    assert np.array_equal(K != 0, g["K"] != 0)


@pytest.mark.parametrize("name", SAL_CASES)
def test_saliency_kernel(golden_dir, name):
    g = load(golden_dir, name)
    dk, dn = rp.saliency_kernel(g["points"], g["normals"], g["K"], nbr_lists(g), float(g["rho"]),
                                float(g["sigma_n"]), float(g["sigma_k"]), float(g["alpha"]),
                                int(g["min_points"]))
    np.testing.assert_allclose(dk, g["dk"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(dn, g["dn"], rtol=1e-6, atol=0)


def test_lattice_ties_included(golden_dir):
    g = load(golden_dir, "lattice_r5")
    pts, q = g["points"], g["queries"]
    want = nbr_lists(g)
    got = rp.query_radius_bruteforce(pts, q, 5.0)
    ties = 0
    for a, b, qq in zip(got, want, q):
        assert np.array_equal(a, b)
        d2 = ((pts[b] - qq) ** 2).sum(1)
        ties += int((d2 == 25.0).sum())
    assert ties > 1000          # exact ties exist and are INCLUDED (<=)
    tree = rp.build_tree(pts)
    assert np.array_equal(np.sort(rp.query_radius(tree, pts[100], 5.0)), g["single"])
    ki, kd = rp.query_knn(tree, q[:50] + 0.25, 5, return_distnaces=True)
    assert np.array_equal(ki, g["knn_i"]) and np.array_equal(kd, g["knn_d"])


def test_lod_cells_match_reference_wrapper(golden_dir):
    g = load(golden_dir, "balltree_6k")
    tree = rp.build_tree(g["points"], 40)
    _, idx_array, nodes, _ = tree.get_arrays()
    assert np.array_equal(idx_array, g["idx_array"])
    assert np.array_equal(nodes["radius"], g["radius"])
    n = int(g["numberOfNodes"])
    # heap layout == the reference's O(n^2) hierarchy reconstruction
    left = np.where(g["is_leaf"] == 1, -1, 2 * np.arange(n) + 1)
    assert np.array_equal(left, g["left"])
    parent = np.where(np.arange(n) == 0, -1, (np.arange(n) - 1) // 2)
    assert np.array_equal(parent, g["parent"])
    for lod in (1.0, 2.0, 5.0):
        assert np.array_equal(rp.lod_cells(tree, lod), g["cells_%g" % lod])


def test_known_answers_plane_and_blend():
    from more3d_b200 import synthetic
    pts = synthetic.plane(1500, 3, normal=(0.1, -0.2, 1.0))
    tree = rp.build_tree(pts)
    nb = rp.query_radius(tree, pts, 0.8)
    nrm, ratio = rp.pca_normals(pts, nb)
    ok = np.array([len(j) >= 3 for j in nb])
    n0 = np.array([0.1, -0.2, 1.0]) / np.linalg.norm([0.1, -0.2, 1.0])
    assert np.all(np.abs(np.abs(nrm[ok] @ n0) - 1) < 1e-12)
    assert np.all(ratio[ok] < 1e-12)
    K, cnt, nrm2 = rp.curvature_kernel(pts, nrm, nb, 0.05, 0.02, 4)
    assert np.all(K == 0)
    dk, dn = rp.saliency_kernel(pts, nrm2, K, nb, 0.8 / np.sqrt(2))
    assert np.all(dk == 0) and np.all(dn == 0)
    sal, mm = rp.saliency_blend(np.array([0, 1, 2], np.float32), np.array([1, 1, 3], np.float32), 0.25)
    # (x - min)/(max - min) + min, reference DM_saliency.py:577-580
    assert np.allclose(sal, 0.75 * np.array([1, 1, 2.0]) + 0.25 * np.array([0, .5, 1]))


def test_otsu_and_voxel_restatements():
    rng = np.random.Generator(np.random.PCG64(1))
    v = np.concatenate([rng.normal(0.2, 0.05, 5000), rng.normal(0.8, 0.05, 3000)])
    t = rp.otsu_threshold(v)
    assert 0.35 < t < 0.65
    pts = rng.random((2000, 3)) * 4
    means, keys = rp.voxel_downsample(pts, 1.0)
    k = rp.voxel_keys(pts, 1.0)
    assert len(means) == len(np.unique(k, axis=0))
    j = np.nonzero(np.all(k == keys[3], axis=1))[0]
    assert np.allclose(means[3], pts[j].mean(0))


def test_levelset_oracle_matches_reference(golden_dir):
    """oracle/ref_levelset.py against the verbatim-reference goldens (frames, WLS operator, 25 iterations)."""
    from oracle import ref_levelset as rl
    g = load(golden_dir, "levelset_4k")
    pts, h, k = g["points"], float(g["h"]), int(g["k"])
    nb, n, t1, t2 = rl.knn_frames(pts, k)
    assert np.array_equal(nb, g["neighbors"])
    np.testing.assert_allclose(n, g["normals"], atol=1e-14)
    np.testing.assert_allclose(t1, g["t1"], atol=1e-13)
    np.testing.assert_allclose(t2, g["t2"], atol=1e-13)
    w, D = rl.wls_operator(pts, nb, g["t1"], g["t2"], 3 * h)
    # numpy's array pow (wendland's **4) is SIMD code whose last bits depend on the host CPU: identical on the
    # machine that generated the golden, within 2 ulp elsewhere.  The run below uses the golden's weights.
    assert np.abs(w.view(np.int64) - g["weights"].view(np.int64)).max() <= 2
    np.testing.assert_allclose(D[:300], g["D"], rtol=1e-9, atol=1e-12)
    w = g["weights"]
    phi0 = rl.init_voxels(pts, 2.0, 2 * h)
    assert np.array_equal(phi0, g["phi0"])
    act0 = rl.zero_crossings(phi0, nb, np.arange(len(pts)))
    assert np.array_equal(act0, g["active0"])
    phi = rl.smooth(phi0, np.arange(len(pts)), nb, w)
    np.testing.assert_allclose(phi, g["phi_s"], rtol=1e-14, atol=1e-15)
    st = dict(points=pts, nbrs=nb, t1=g["t1"], t2=g["t2"], w=w, D=D, h=h, zeta=g["zeta"], phi=phi, active=act0,
              last_active=act0, alias=True)
    kw = dict(stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
    assert rl.run(st, 10, **kw) is False
    assert np.array_equal(st["active"], g["act10"])
    np.testing.assert_allclose(st["phi"], g["phi10"], rtol=1e-9, atol=1e-11)
    rl.run(st, 15, **kw)
    assert np.array_equal(st["active"], g["act25"])
    np.testing.assert_allclose(st["phi"], g["phi25"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("tag,kw", [("a", dict(nu=1e-4, mu=2.5e-3, cap=True)), ("b", dict(nu=0.0, mu=0.0, cap=True)),
                                    ("c", dict(nu=2e-4, mu=0.0, cap=False))])
def test_levelset_oracle_mask_init(golden_dir, tag, kw):
    from oracle import ref_levelset as rl
    g = load(golden_dir, "levelset_mask_3k")
    pts, h = g["points"], float(g["h"])
    w, D = rl.wls_operator(pts, g["neighbors"], g["t1"], g["t2"], 3 * h)
    phi = np.where(g["mask"], -2 * h, 2 * h)
    act = rl.zero_crossings(phi, g["neighbors"], np.arange(len(pts)))
    st = dict(points=pts, nbrs=g["neighbors"], t1=g["t1"], t2=g["t2"], w=w, D=D, h=h, zeta=g["zeta"], phi=phi,
              active=act, last_active=act.copy(), alias=False)
    rl.run(st, 12, stepsize=.01, lambda1=1.5, lambda2=0.5, epsilon=0.8, tolerance=0, **kw)
    assert np.array_equal(st["active"], g["act_" + tag]) and np.array_equal(st["last_active"], g["last_" + tag])
    np.testing.assert_allclose(st["phi"], g["phi_" + tag], rtol=1e-9, atol=1e-11)


LMV_RUNS = {"v": dict(stepsize=.005, nu=1e-4, mu=2.5e-3, epsilon=1, cap=True),
            "m": dict(stepsize=.01, nu=2e-4, mu=1e-3, epsilon=0.8, cap=True)}


@pytest.mark.parametrize("tag", ["v", "m"])
def test_levelset_oracle_lmv(golden_dir, tag):
    """model='lmv' (_lmv_step): oracle vs the verbatim reference run with the caller-squeezed 1-D cue
    (tools/make_golden.py levelset_lmv); the golden also records that the (N, 1) cue raises in the reference."""
    from oracle import ref_levelset as rl
    g = load(golden_dir, "levelset_lmv_3k")
    assert "could not be broadcast" in str(g["raised_on_2d_cue"])
    pts, h = g["points"], float(g["h"])
    _, D = rl.wls_operator(pts, g["neighbors"], g["t1"], g["t2"], 3 * h)
    w = g["weights"]
    if tag == "v":
        phi = rl.init_voxels(pts, 2.0, 2 * h)
        alias = True
    else:
        phi = np.where((pts[:, 0] ** 2 + pts[:, 1] ** 2) < 9.0, -2 * h, 2 * h)
        alias = False
    act = rl.zero_crossings(phi, g["neighbors"], np.arange(len(pts)))
    st = dict(points=pts, nbrs=g["neighbors"], t1=g["t1"], t2=g["t2"], w=w, D=D, h=h, zeta=g["zeta"], phi=phi,
              active=act, last_active=act if alias else act.copy(), alias=alias)
    kw = dict(LMV_RUNS[tag], lambda1=1, lambda2=1, tolerance=0, model="lmv")
    rl.run(st, 1, **kw)
    assert np.array_equal(st["active"], g["act1_" + tag])
    np.testing.assert_allclose(st["phi"], g["phi1_" + tag], rtol=1e-9, atol=1e-11)
    rl.run(st, 11, **kw)
    assert np.array_equal(st["active"], g["act12_" + tag])
    np.testing.assert_allclose(st["phi"], g["phi12_" + tag], rtol=1e-9, atol=1e-11)
    if not alias:
        assert np.array_equal(st["last_active"], g["last12_" + tag])


def test_verbatim_reference_bytecode_reproduces_the_goldens(golden_dir):
    """oracle/_ref (the UNMODIFIED reference byte-compiled by oracle/build_ref.py) is what the GPU-box tests and
    bench.py's reference arm execute: it must reproduce the goldens that were generated from the sources."""
    from oracle import ref_verbatim as rv
    if not rv.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    g = np.load(os.path.join(golden_dir, "terrain_r05.npz"))
    pts = g["points"]
    off, idx = g["nbr_off"], g["nbr_idx"]
    nb = [idx[off[i]:off[i + 1]].astype(np.int64) for i in range(len(off) - 1)]
    sub = np.arange(0, len(pts), 3)
    K, pc, nrm = rv.run_curvature(pts, g["normals_in"][sub], [nb[i] for i in sub], float(g["alpha"]), float(g["roughness"]),
                                  int(g["min_points"]), queries=sub)
    assert np.array_equal(K, g["K"][sub]) and np.array_equal(pc, g["pcount"][sub])
    assert np.array_equal(nrm, g["normals"][sub], equal_nan=True)
    dk, dn = rv.run_dk_dn(pts, g["normals"], g["K"], [nb[i] for i in sub], float(g["rho"]), float(g["rho"]), float(g["sigma_n"]),
                          float(g["sigma_k"]), float(g["alpha"]), int(g["min_points"]), queries=sub)
    assert np.array_equal(dk, g["dk"][sub]) and np.array_equal(dn, g["dn"][sub])
    bt_mod, _, ls_mod = rv.load_reference()
    bt = bt_mod.BallTreePointSet(pts[:500], leaf_size=40)
    assert bt.numberOfNodes >= 1 and hasattr(ls_mod, "Solver")


def test_cpu_bench_kinds_agree():
    """bench.py's CPU legs: the verbatim reference and the oracle port time the same work (same outputs)."""
    from more3d_b200 import synthetic
    from oracle import cpu_bench, ref_verbatim as rv
    pts = synthetic.terrain(20000, seed=4)
    P = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
    kinds = ["port"] + (["reference"] if rv.available() else [])
    outs = {}
    for kind in kinds:
        b = cpu_bench.CpuSaliencyBaseline(pts, (0.5,), P, workers=1, seed=1, kind=kind)
        sub = np.arange(0, 20000, 200)
        outs[kind] = (cpu_bench._pass1_chunk((0.5, sub)), cpu_bench._pass2_chunk((0.5, sub)))
        total, raw = b.step(50)
        assert total > raw > 0
    if "reference" in outs:
        np.testing.assert_allclose(outs["port"][0], outs["reference"][0], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(outs["port"][1], outs["reference"][1], rtol=1e-4, atol=1e-7)


def test_otsu_restatement_against_opencv():
    """skimage (the reference's `threshold_otsu`, Downsample.py:54) is absent offline, so the restatement in
    oracle/ref_path.py cannot be pinned on skimage itself.  An INDEPENDENT implementation of the same criterion is here:
    OpenCV's THRESH_OTSU on 8-bit data.  On integer-valued data with both ends present every value has its own bin in the
    float path (256 bins over [0, 255]), the returned bin centre identifies the split 'bins <= idx | bins > idx', and that
    split must be OpenCV's 'v <= T | v > T' (maximum between-class variance, first maximum)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import ref_path as rp
    rng = np.random.Generator(np.random.PCG64(1))
    for trial in range(120):
        n = 20000
        a = rng.normal(rng.uniform(30, 120), rng.uniform(5, 30), n // 2)
        b = rng.normal(rng.uniform(130, 230), rng.uniform(5, 30), n - n // 2)
        img = np.clip(np.rint(np.concatenate([a, b])), 0, 255).astype(np.uint8)
        img[0], img[1] = 0, 255
        T, _ = cv2.threshold(img.reshape(1, -1), 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        t = rp.otsu_threshold(img.astype(np.float64))
        idx = int(round(t * 256 / 255 - 0.5))              # bin centre (idx + 0.5) * 255 / 256 -> idx
        assert idx == int(T), (trial, T, idx, t)
