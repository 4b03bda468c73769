"""Generate tests/golden/*.npz by executing the reference's own code (container only).

    python tools/make_golden.py            # rewrites every fixture

Every array below is the output of /root/reference code run verbatim through
tools/ref_harness.py, except the PCA normals fed into the curvature pass: the reference takes
them as a precondition (OPALS Normals module, DM_saliency.py:329-330), so they come from
oracle.ref_path.pca_normals and are stored in the fixture as an *input*.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_harness  # noqa: E402
from more3d_b200 import synthetic  # noqa: E402
from oracle import ref_path  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def csr(nbrs):
    off = np.zeros(len(nbrs) + 1, np.int64)
    off[1:] = np.cumsum([len(j) for j in nbrs])
    idx = np.concatenate([np.sort(j) for j in nbrs]).astype(np.int32)
    return off, idx


def saliency_case(name, pts, r, sigma_n, sigma_k, roughness=0.02, alpha=0.05, min_points=4):
    bt_mod, _, _ = ref_harness.load_reference()
    bt = bt_mod.BallTreePointSet(pts, leaf_size=40)
    nbrs = bt.queryRadius(pts, r)
    normals_in, ratio = ref_path.pca_normals(pts, nbrs)
    K, pcount, normals = ref_harness.run_curvature(pts, normals_in, nbrs, alpha, roughness, min_points)
    rho = r / np.sqrt(2.0)
    dk, dn = ref_harness.run_dk_dn(pts, normals, K, nbrs, rho, rho, sigma_n, sigma_k, alpha, min_points)
    off, idx = csr(nbrs)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), points=pts, radius=r, rho=rho, sigma_n=sigma_n,
        sigma_k=sigma_k, roughness=roughness, alpha=alpha, min_points=min_points,
        nbr_off=off, nbr_idx=idx, normals_in=normals_in, lam_ratio=ratio, K=K, pcount=pcount,
        normals=normals, dk=dk, dn=dn)
    print(name, "N", len(pts), "mean m %.1f" % (len(idx) / len(pts)),
          "nonzero K %.3f dk %.3f dn %.3f" % ((K != 0).mean(), (dk != 0).mean(), (dn != 0).mean()))


def lattice_case():
    bt_mod, _, _ = ref_harness.load_reference()
    pts = synthetic.lattice(13, 13, 13)
    bt = bt_mod.BallTreePointSet(pts, leaf_size=40)
    q = pts[::7]
    nbrs = bt.queryRadius(q, 5.0)
    off, idx = csr(nbrs)
    one = bt.queryRadius(pts[100], 5.0)         # 1-D input path, BallTreePointSet.py:338-342
    knn_i, knn_d = bt.query(q[:50] + 0.25, 5, return_distnaces=True)
    np.savez_compressed(os.path.join(OUT, "lattice_r5.npz"), points=pts, queries=q, radius=5.0,
                        nbr_off=off, nbr_idx=idx, single=np.sort(one), knn_i=knn_i, knn_d=knn_d)
    print("lattice pairs", len(idx))


def balltree_case():
    """Node structure of the reference wrapper on a small terrain (LOD cells, hierarchy)."""
    bt_mod, _, _ = ref_harness.load_reference()
    pts = synthetic.terrain(6000, seed=21)
    bt = bt_mod.BallTreePointSet(pts, leaf_size=40)
    data = bt.data.get_arrays()
    nodes = data[2]
    out = dict(points=pts, idx_array=np.asarray(data[1]), idx_start=nodes["idx_start"],
               idx_end=nodes["idx_end"], is_leaf=nodes["is_leaf"], radius=nodes["radius"],
               numberOfNodes=bt.numberOfNodes, maxLevel=bt.maxLevel, leaves=bt.ballTreeLeaves,
               levels=np.array([bt.getNodeLevel(i) for i in range(bt.numberOfNodes)]))
    for lod in (1.0, 2.0, 5.0):
        cells = np.atleast_1d(bt.getSmallestNodesOfSize(lod))
        out["cells_%g" % lod] = cells
    out["left"] = np.array([-1 if bt.getLeftChildOfNode(i) is None else bt.getLeftChildOfNode(i)
                            for i in range(bt.numberOfNodes)])
    out["parent"] = np.array([-1 if bt.getParentOfNode(i) is None else bt.getParentOfNode(i)
                              for i in range(bt.numberOfNodes)])
    out["sibling"] = np.array([-1 if bt.getSiblingOfNode(i) is None else bt.getSiblingOfNode(i)
                               for i in range(bt.numberOfNodes)])
    np.savez_compressed(os.path.join(OUT, "balltree_6k.npz"), **out)
    print("balltree nodes", bt.numberOfNodes, "maxLevel", bt.maxLevel)


def levelset_case():
    """build_neighborhoods + Solver.run, verbatim (levelsets_func.py:154-172, 361-642)."""
    _, _, ls = ref_harness.load_reference()
    pts = synthetic.terrain(4000, seed=31)
    pts = pts - pts.mean(axis=0)                  # read_las centres the cloud, levelsets_func.py:63-64
    h, k = 0.5, 7
    neighbors, normals, tangents = ls.build_neighborhoods(pts.copy(), h, k)
    rng = np.random.Generator(np.random.PCG64(5))
    zeta = np.abs(np.sin(pts[:, :1] / 3.0) * np.cos(pts[:, 1:2] / 2.0)) + 0.05 * rng.random((len(pts), 1))
    solver = ls.Solver(h, zeta, pts, neighbors, tangents, "chan_vese")
    z = solver.zeta
    ls.normalize(z)
    solver.initialize_from_voxels(2.0)
    phi0 = solver.phi.copy()
    active0 = solver.active.copy()
    solver.phi[:] = ls.smooth(solver.phi, slice(None), neighbors, solver.diff.weights)
    phi_s = solver.phi.copy()
    dx, dy, ddx, ddy = solver.diff(solver.phi, slice(None))
    conv = solver.run(10, stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1,
                      cap=True, tolerance=0)
    phi10 = solver.phi.copy()
    act10 = solver.active.copy()
    solver.run(15, stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1,
               cap=True, tolerance=0)
    np.savez_compressed(
        os.path.join(OUT, "levelset_4k.npz"), points=pts, h=h, k=k, neighbors=neighbors,
        normals=normals, t1=tangents[0], t2=tangents[1], zeta=solver.zeta, weights=solver.diff.weights,
        D=solver.diff.D[:300], phi0=phi0, active0=active0, phi_s=phi_s, dx=dx, dy=dy, ddx=ddx, ddy=ddy,
        phi10=phi10, act10=act10, phi25=solver.phi, act25=solver.active, conv=conv)
    print("levelset active", act10.sum(), solver.active.sum())


def levelset_mask_case():
    """initialize_from_mask (separate active / last_active -> the re-clamp of (de)activated points runs),
    two cue channels, k = 5, no curvature term in one run and no cap in another."""
    _, _, ls = ref_harness.load_reference()
    pts = synthetic.terrain(3000, seed=33, noise=0.05)
    pts = pts - pts.mean(axis=0)
    h, k = 0.6, 5
    neighbors, normals, tangents = ls.build_neighborhoods(pts.copy(), h, k)
    rng = np.random.Generator(np.random.PCG64(9))
    zeta = np.column_stack([np.abs(np.sin(pts[:, 0] / 2.0)), rng.random(len(pts)) ** 3])
    out = dict(points=pts, h=h, k=k, neighbors=neighbors, normals=normals, t1=tangents[0], t2=tangents[1], zeta=zeta)
    mask = (pts[:, 0] ** 2 + pts[:, 1] ** 2) < 9.0
    for tag, kw in (("a", dict(nu=1e-4, mu=2.5e-3, cap=True)), ("b", dict(nu=0.0, mu=0.0, cap=True)),
                    ("c", dict(nu=2e-4, mu=0.0, cap=False))):
        solver = ls.Solver(h, zeta, pts, neighbors, tangents, "chan_vese")
        solver.initialize_from_mask(mask)
        solver.run(12, stepsize=.01, lambda1=1.5, lambda2=0.5, epsilon=0.8, tolerance=0, **kw)
        out["phi_" + tag] = solver.phi.copy()
        out["act_" + tag] = solver.active.copy()
        out["last_" + tag] = solver.last_active.copy()
    out["mask"] = mask
    np.savez_compressed(os.path.join(OUT, "levelset_mask_3k.npz"), **out)
    print("levelset mask case", [int(out["act_" + t].sum()) for t in "abc"])


def levelset_lmv_case():
    """Solver.run with active_contour_model='lmv' (_lmv_step, levelsets_func.py:439-513), executed verbatim.
    With the (N, C) cue the constructor requires, the step raises ValueError (zeta[nb] * (1 - Hphi[nb]) cannot
    broadcast); it is written for a 1-D cue and runs unmodified once the CALLER squeezes the attribute --
    that one assignment is the only thing this harness adds."""
    _, _, ls = ref_harness.load_reference()
    pts = synthetic.terrain(3000, seed=41, noise=0.03)
    pts = pts - pts.mean(axis=0)
    h, k = 0.5, 7
    neighbors, normals, tangents = ls.build_neighborhoods(pts.copy(), h, k)
    rng = np.random.Generator(np.random.PCG64(13))
    zeta = np.abs(np.sin(pts[:, :1] / 3.0) * np.cos(pts[:, 1:2] / 2.0)) + 0.1 * rng.random((len(pts), 1))
    out = dict(points=pts, h=h, k=k, neighbors=neighbors, normals=normals, t1=tangents[0], t2=tangents[1], zeta=zeta)
    s0 = ls.Solver(h, zeta, pts, neighbors, tangents, "lmv")
    s0.initialize_from_voxels(2.0)
    try:
        s0.run(1, stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
        raised = ""
    except ValueError as e:
        raised = str(e)
    assert raised, "the reference's lmv step was expected to raise on an (N, 1) cue"
    out["raised_on_2d_cue"] = raised
    for tag, init, kw in (("v", "voxels", dict(stepsize=.005, nu=1e-4, mu=2.5e-3, epsilon=1, cap=True)),
                          ("m", "mask", dict(stepsize=.01, nu=2e-4, mu=1e-3, epsilon=0.8, cap=True))):
        solver = ls.Solver(h, zeta, pts, neighbors, tangents, "lmv")
        solver.zeta = solver.zeta[:, 0]                 # the caller-side squeeze
        if init == "voxels":
            solver.initialize_from_voxels(2.0)
        else:
            solver.initialize_from_mask((pts[:, 0] ** 2 + pts[:, 1] ** 2) < 9.0)
        out["weights"] = solver.diff.weights
        solver.run(1, lambda1=1, lambda2=1, tolerance=0, **kw)
        out["phi1_" + tag] = solver.phi.copy()
        out["act1_" + tag] = solver.active.copy()
        solver.run(11, lambda1=1, lambda2=1, tolerance=0, **kw)
        out["phi12_" + tag] = solver.phi.copy()
        out["act12_" + tag] = solver.active.copy()
        out["last12_" + tag] = solver.last_active.copy()
    np.savez_compressed(os.path.join(OUT, "levelset_lmv_3k.npz"), **out)
    print("levelset lmv case", int(out["act12_v"].sum()), int(out["act12_m"].sum()), "|", raised)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["sal", "lattice", "balltree", "levelset", "levelset_mask", "levelset_lmv"]
    if "sal" in which:
        saliency_case("terrain_r05", synthetic.terrain(3000, seed=11), 0.5, 0.01, 0.01)
        saliency_case("terrain_r10", synthetic.terrain(2500, seed=12), 1.0, 0.01, 0.01)
        saliency_case("rough_r05", synthetic.terrain(3000, seed=13, noise=0.15), 0.5, 1e-3, 0.01)
        saliency_case("boulders_r075", synthetic.terrain(2500, seed=14, noise=0.05, boulders=40), 0.75, 1e-3, 0.01)
        # coincident points (np.unique path, DM_saliency.py:247) and sparse fringes (m < min_points, NaN normals)
        base = synthetic.terrain(1800, seed=15, noise=0.12)
        rng = np.random.Generator(np.random.PCG64(15))
        dup = rng.choice(len(base), 250, replace=False)
        far = base[:40] + rng.normal(0, 6.0, (40, 3)) + [0, 0, 30.0]
        pts = np.vstack([base, base[dup], base[dup[:50]], far])
        saliency_case("dups_sparse_r06", pts[rng.permutation(len(pts))], 0.6, 1e-3, 0.01)
    if "lattice" in which:
        lattice_case()
    if "balltree" in which:
        balltree_case()
    if "levelset" in which:
        levelset_case()
    if "levelset_mask" in which:
        levelset_mask_case()
    if "levelset_lmv" in which:
        levelset_lmv_case()
