"""CPU ORACLE -- test infrastructure, NOT product code.

A numpy / scikit-learn restatement of the reference's per-point neighbourhood-feature path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module; the product package ``more3d_b200`` never does.

Pinning status: the reference ships no tests or golden vectors, so the pins are
(1) ``tests/golden/*.npz`` produced by executing the reference's own code verbatim in the build
container (``tools/make_golden.py`` -> ``tools/ref_harness.py``) and checked against this module by
``tests/test_oracle_golden.py``; (2) analytic known answers (plane, sphere, lattice ties).
The pieces the reference delegates to packages that are absent offline -- OPALS Histo/AddInfo
(``saliency_blend``), scikit-image ``threshold_otsu`` (``otsu_threshold``), open3d
``voxel_down_sample`` (``voxel_downsample``) and the not-executable-as-shipped loop of
Downsample.py:44-83 (``saliency_downsample``, a declared repair) -- are restated from their
published algorithms: **parity unpinned** for those four.

Third-party arithmetic at the boundary: scikit-learn 1.9.0 ``BallTree`` (reference pins no version;
call sites BallTreePointSet.py:35, 311, 339, 345; levelsets_func.py:158-163).
"""
import numpy as np
from scipy import stats
from sklearn.neighbors import BallTree


# --------------------------------------------------------------------------------------------
# H1/H2/H3  BallTreePointSet.__init__/queryRadius/query  (BallTreePointSet.py:19-53, 292-347)
# --------------------------------------------------------------------------------------------
def build_tree(points, leaf_size=40):
    """BallTreePointSet.py:35 -- sklearn BallTree over float64 C-order points."""
    return BallTree(np.ascontiguousarray(points, dtype=np.float64), leaf_size)


def query_radius(tree, pnts, radius):
    """BallTreePointSet.py:317-347: object array of int64 index arrays (unsorted, self included)."""
    pnts = np.asarray(pnts, dtype=np.float64)
    if pnts.ndim == 1:
        return tree.query_radius(pnts.reshape(1, -1), radius)[0]
    return tree.query_radius(pnts, radius)


def query_radius_bruteforce(points, pnts, radius):
    """The accept rule the ball tree implements (sklearn _dist_metrics.pxd.tp:39-49 +
    _binary_tree.pxi.tp:1951-1957): fp64, ((dx*dx)+(dy*dy))+(dz*dz) <= r*r, left to right, no FMA."""
    points = np.asarray(points, dtype=np.float64)
    pnts = np.atleast_2d(np.asarray(pnts, dtype=np.float64))
    r2 = np.float64(radius) * np.float64(radius)
    out = np.empty(pnts.shape[0], dtype=object)
    for i, q in enumerate(pnts):
        dx = q[0] - points[:, 0]
        dy = q[1] - points[:, 1]
        dz = q[2] - points[:, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        out[i] = np.nonzero(d2 <= r2)[0].astype(np.int64)
    return out


def query_knn(tree, pnts, k, return_distnaces=False):
    """BallTreePointSet.py:292-315 (keyword spelling as in the reference)."""
    d, i = tree.query(np.atleast_2d(np.asarray(pnts, dtype=np.float64)), k=k)
    return (i, d) if return_distnaces else i


# --------------------------------------------------------------------------------------------
# H4b  centroid-PCA normals + surface variation  (named by north_star; downsample_compare.py:53-64)
# --------------------------------------------------------------------------------------------
def pca_point(p, q):
    """PCA of one neighbourhood q (m,3) of the point p.  Returns (normal, ratio); m < 3 -> NaN."""
    m = q.shape[0]
    if m < 3:
        return np.full(3, np.nan), 0.0
    d = q - p
    c = d.mean(axis=0)
    e = d - c
    cov = e.T @ e / m
    w, v = np.linalg.eigh(cov)
    s = w.sum()
    return v[:, 0], (w[0] / s if s > 0 else 0.0)


def pca_normals(points, nbrs):
    """Unit eigenvector of the smallest eigenvalue of the centroid-centred covariance of each
    neighbourhood, and lambda_min / sum(lambda).  Sign is arbitrary.  m < 3 -> NaN normal."""
    n = points.shape[0]
    normals = np.full((n, 3), np.nan)
    ratio = np.zeros(n)
    for i in range(n):
        normals[i], ratio[i] = pca_point(points[i], points[nbrs[i]])
    return normals, ratio


# --------------------------------------------------------------------------------------------
# H5  Curvature_Kernel.process  (DM_saliency.py:88-157)
# --------------------------------------------------------------------------------------------
def curvature_epsilon(alpha, roughness):
    """DM_saliency.py:99."""
    return float(stats.norm.ppf(1 - alpha / 2) * roughness)


def curvature_point(p, normal, q, eps, min_points):
    """One point of Curvature_Kernel.process.  Returns (K, pcount, normal_out)."""
    m = q.shape[0]
    if np.isnan(normal[0]):                                   # :107-110
        return 0.0, m, normal
    if m < min_points:                                        # :120-123
        return 0.0, m, normal
    proj_mat = np.eye(3) - np.outer(normal, normal)           # :117 (before the flip)
    if normal.dot(-p) < 0:                                    # :131-135
        normal = -normal
    cg = (proj_mat @ q.T).T - proj_mat @ p                    # :139-141
    cg = cg.mean(axis=0)
    if np.linalg.norm(cg) > 1:                                # :142-144
        return 0.0, m, normal
    with np.errstate(all="ignore"):
        sum_proj = np.sum(normal @ (p - q).T) / (m - 1)       # :137-146
    if abs(sum_proj) < eps:                                   # :149-153
        return 0.0, m, normal
    return float(sum_proj), m, normal


def curvature_kernel(points, normals, nbrs, alpha, roughness, min_points=4):
    """Loop of ProcessorEx.run(Curvature_Kernel) (DM_saliency.py:403-407).
    Returns (_nonparam_K float32, _pcount int32, normals float64 possibly flipped)."""
    n = points.shape[0]
    eps = curvature_epsilon(alpha, roughness)
    K = np.zeros(n, np.float32)          # OPALS ColumnType.float_ (:381)
    cnt = np.zeros(n, np.int32)          # :382
    out_n = np.array(normals, dtype=np.float64, copy=True)
    for i in range(n):
        k, m, nn = curvature_point(points[i], out_n[i], points[nbrs[i]], eps, min_points)
        K[i] = k
        cnt[i] = m
        out_n[i] = nn
    return K, cnt, out_n


# --------------------------------------------------------------------------------------------
# H6  Saliency_Kernel.process  (DM_saliency.py:219-324)
# --------------------------------------------------------------------------------------------
def chi2_table(alpha, max_m):
    """chi2.ppf(alpha, df) for df = 0..max_m (DM_saliency.py:282); df = 0 -> NaN."""
    with np.errstate(all="ignore"):
        return stats.chi2.ppf(alpha, np.arange(max_m + 1)).astype(np.float64)


def saliency_point(p, normal, Kp, q, qn, qK, rho, sigma_normal, sigma_curvature, table, min_points):
    """One point of Saliency_Kernel.process.  Returns (dk, dn)."""
    m = q.shape[0]
    if np.isnan(normal[0]) or m < min_points:                 # :225-229, :240-244
        return 0.0, 0.0
    _, first = np.unique(q.T, axis=1, return_index=True)      # :247 dedup coincident nbrs
    q, qn, qK = q[first], qn[first], qK[first]
    d = p - q
    dist = np.sum(d ** 2, axis=1)                             # squared distance, :257
    with np.errstate(all="ignore"):
        pn = normal / np.linalg.norm(normal)                  # :258
        qn = qn / np.linalg.norm(qn, axis=1)[:, None]         # :259
        dn = 1 - qn @ pn                                      # :261
        dk = Kp - qK                                          # :262
        rho2 = rho ** 2
        w = np.zeros(dist.shape[0])                           # :272-275
        lo = dist < rho2
        hi = dist > rho2
        w[lo] = 1 / rho2 * dist[lo]
        w[hi] = -1 / rho2 * dist[hi] + 2
        w[w < 0] = 0
        act = w > 0.01                                        # :279
        n_act = int(act.sum())
        chi_t = table[n_act]                                  # :282
        chi_n = (m - 1) * dn.std() / sigma_normal             # :283
        chi_k = (m - 1) * dk.std() / sigma_curvature          # :284
        z_n = int(np.sum(np.abs(dn[act]) < 0.016))            # :286
        z_k = int(np.sum(np.abs(dk[act]) <= 0.04))            # :287
        if z_k == n_act:                                      # :289-290
            z_k = int(np.sum(np.abs(dk[act] * 100) <= 0.04))
        if chi_n < chi_t or z_n > 0.6 * n_act:                # :292-300
            dn_out = 0.0
        else:
            dn_out = np.sum(np.abs(dn) * w) / np.sum(w)
        if np.isnan(dn_out):                                  # :303-308
            dn_out = 0.0
        elif dn_out > 2:
            dn_out = 1.0
        if chi_k < chi_t or z_k > 0.6 * n_act:                # :309-316
            dk_out = 0.0
        else:
            dk_out = np.abs(np.sum(dk * w) / np.sum(w))
    return float(dk_out), float(dn_out)


def saliency_kernel(points, normals, K, nbrs, rho, sigma_normal=0.01, sigma_curvature=0.1,
                    alpha=0.05, min_points=4):
    """Loop of ProcessorEx.run(Saliency_Kernel) (DM_saliency.py:516-522).
    Returns (_dk float32, _dn float32)  (float_ columns, :491-492)."""
    n = points.shape[0]
    max_m = max((len(j) for j in nbrs), default=0)
    table = chi2_table(alpha, max_m)
    K64 = np.asarray(K, np.float32).astype(np.float64)
    dk = np.zeros(n, np.float32)
    dn = np.zeros(n, np.float32)
    for i in range(n):
        j = nbrs[i]
        a, b = saliency_point(points[i], normals[i], K64[i], points[j], normals[j], K64[j],
                              rho, sigma_normal, sigma_curvature, table, min_points)
        dk[i] = a
        dn[i] = b
    return dk, dn


# --------------------------------------------------------------------------------------------
# H7  DM_saliency  (DM_saliency.py:562-580) -- OPALS Histo/AddInfo restated: parity unpinned
# --------------------------------------------------------------------------------------------
def saliency_blend(dk, dn, curvature_weight=0.5):
    """min/max normalise (with the reference's '+ min' term) and blend.  The f32 attribute
    columns are evaluated in fp64 and stored back as f32 (AddInfo '(float)' columns)."""
    dk = np.asarray(dk, np.float32)
    dn = np.asarray(dn, np.float32)
    min_k, max_k = float(dk.min()), float(dk.max())           # :565-568
    min_n, max_n = float(dn.min()), float(dn.max())
    diff_k, diff_n = max_k - min_k, max_n - min_n             # :570-571
    nw = 1 - curvature_weight                                 # :572
    with np.errstate(all="ignore"):
        ndk = ((dk.astype(np.float64) - min_k) / diff_k + min_k).astype(np.float32)   # :577
        ndn = ((dn.astype(np.float64) - min_n) / diff_n + min_n).astype(np.float32)   # :578
        sal = (nw * ndn.astype(np.float64) + curvature_weight * ndk.astype(np.float64)
               ).astype(np.float32)                                                  # :579-580
    return sal, (min_k, max_k, min_n, max_n)


def saliency_stats(sal):
    """DM_saliency_stats (DM_saliency.py:585-607): (min, max, mean, sigma)."""
    s = np.asarray(sal, np.float64)
    return float(s.min()), float(s.max()), float(s.mean()), float(s.std())


# --------------------------------------------------------------------------------------------
# H8  Downsample.py:44-83 (declared repair) + threshold_otsu -- parity unpinned (skimage absent; the Otsu criterion is
#     cross-checked against OpenCV on 8-bit data, tests/test_oracle_golden.py::test_otsu_restatement_against_opencv)
# --------------------------------------------------------------------------------------------
def otsu_histogram(values, nbins=256):
    v = np.asarray(values, np.float64).ravel()
    counts, edges = np.histogram(v, bins=nbins)
    centers = (edges[:-1] + edges[1:]) / 2.0
    return counts, centers, edges


def otsu_from_histogram(counts, centers):
    """skimage.filters.threshold_otsu (0.19.x) on a precomputed histogram."""
    counts = counts.astype(np.float64)
    with np.errstate(all="ignore"):
        w1 = np.cumsum(counts)
        w2 = np.cumsum(counts[::-1])[::-1]
        m1 = np.cumsum(counts * centers) / w1
        m2 = (np.cumsum((counts * centers)[::-1]) / w2[::-1])[::-1]
        var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return float(centers[int(np.argmax(var12))])


def otsu_threshold(values, nbins=256):
    """Downsample.py:54."""
    v = np.asarray(values, np.float64).ravel()
    if np.all(v == v[0]):
        return float(v[0])
    counts, centers, _ = otsu_histogram(v, nbins)
    return otsu_from_histogram(counts, centers)


def lod_cells(tree, radius):
    """getSmallestNodesOfSize(radius, 'root') (BallTreePointSet.py:128-147, 256-271): first node on
    each branch with node radius < LOD or leaf; children of node i are 2i+1, 2i+2 (heap layout of
    sklearn's tree, verified equal to __findChildNodes by the survey probe).  Returns the node
    ids in the reference's depth-first (left before right) order."""
    _, idx_array, nodes, _ = tree.get_arrays()
    rad = nodes["radius"]
    leaf = nodes["is_leaf"]
    out = []
    stack = [0]
    while stack:
        i = stack.pop()
        if rad[i] < radius or leaf[i] == 1:
            out.append(i)
        else:
            stack.append(2 * i + 2)
            stack.append(2 * i + 1)
    return np.asarray(out, dtype=np.int64)


def points_of_node(tree, i):
    """getPointsOfNode (BallTreePointSet.py:240-254)."""
    _, idx_array, nodes, _ = tree.get_arrays()
    return np.asarray(idx_array[nodes["idx_start"][i]:nodes["idx_end"][i]])


def saliency_downsample(points, saliency, lod, leaf_size=40, literal=False, zero_thresh=1e-4, tree=None):
    """Downsample.py:44-83 with the declared repair of :58/:67:
    ``idx_saliency = saliency > s_thresh``; non-salient mean over ``~idx_saliency``.
    ``literal=True`` keeps both conjuncts of :58 with ``and`` -> ``&``.
    Returns (kept_xyz (M,3), flag (M,), per-cell list of (kept original indices, flags))."""
    points = np.ascontiguousarray(points, np.float64)
    saliency = np.asarray(saliency, np.float64)
    if tree is None:
        tree = build_tree(points, leaf_size)
    s_thresh = otsu_threshold(saliency)                        # :54
    cells = lod_cells(tree, lod)                               # :47
    kept_idx, kept_flag = [], []
    for c in cells:
        node_idx = points_of_node(tree, c)                     # :57
        sal = saliency[node_idx]
        mask = sal > s_thresh
        if literal:
            mask = mask & (np.abs(sal) < zero_thresh)
        if mask.sum() * 100 / mask.shape[0] > 60:              # :60-63
            kept_idx.append(node_idx)
            kept_flag.append(np.ones(node_idx.shape[0]))
        else:
            avg = points[node_idx][~mask].mean(axis=0)         # :67 (repaired complement)
            near = tree.query(avg[None, :], k=1, return_distance=False)[0][0]   # :70
            kept_idx.append(np.concatenate([node_idx[mask], [near]]))           # :71
            f = np.zeros(int(mask.sum()) + 1)
            f[:int(mask.sum())] = 1                            # :74-75
            kept_flag.append(f)
    idx = np.concatenate(kept_idx) if kept_idx else np.zeros(0, np.int64)
    flag = np.concatenate(kept_flag) if kept_flag else np.zeros(0)
    return points[idx], flag, idx, cells


# --------------------------------------------------------------------------------------------
# H9  open3d voxel_down_sample (Downsample.py:86-89) -- parity unpinned
# --------------------------------------------------------------------------------------------
def voxel_keys(points, voxel_size):
    points = np.asarray(points, np.float64)
    vmin = points.min(axis=0) - voxel_size * 0.5
    return np.floor((points - vmin) / voxel_size).astype(np.int64)


def voxel_downsample(points, voxel_size):
    """Mean of the member points of every occupied voxel; members are accumulated in original
    point order (open3d AccumulatedPoint::AddPoint).  Returns (means (V,3), keys (V,3) int64)
    sorted lexicographically by voxel key (open3d's own order is hash-map order)."""
    points = np.asarray(points, np.float64)
    keys = voxel_keys(points, voxel_size)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))   # stable: keeps original order inside a voxel
    ks = keys[order]
    new = np.ones(len(order), bool)
    new[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    starts = np.nonzero(new)[0]
    ends = np.append(starts[1:], len(order))
    means = np.empty((len(starts), 3))
    for v, (a, b) in enumerate(zip(starts, ends)):
        acc = np.zeros(3)
        for j in order[a:b]:
            acc += points[j]
        means[v] = acc / float(b - a)
    return means, ks[starts]


# --------------------------------------------------------------------------------------------
# "next" row f-2: downsample_compare.py:132-150 (error terms for given normals)
# --------------------------------------------------------------------------------------------
def compare_downsampled(pts, downsampled, normals_pts, normals_down, leafsize=40):
    """Literal restatement: c2c = avg(E2) + avg(E)/2 (:142-143), c2p with the original normals indexed by
    closest2 (:145), dN = median folded angle (:146-150)."""
    pts = np.asarray(pts, np.float64)[:, :3]
    down = np.asarray(downsampled, np.float64)[:, :3]
    btP = BallTree(pts, leafsize)
    dbt = BallTree(down, leafsize)
    E, closest = btP.query(down, 1)
    closest = closest.flatten()
    E2, closest2 = dbt.query(pts, 1)
    closest2 = closest2.flatten()
    vector_E2 = pts - down[closest2]
    c2c = np.average(E2) + np.average(E) / 2
    c2p = np.average(np.einsum('ij,ij->i', vector_E2, normals_pts[closest2, :]) ** 2)
    dn = np.einsum('ij,ij->i', normals_pts[closest, :], normals_down)
    dn[np.abs(dn) > 1] = 1
    dn = np.degrees(np.arccos(dn))
    dn[dn > 90] = 180 - dn[dn > 90]
    return {"c2c": float(c2c), "c2p": float(c2p), "dN": float(np.median(dn)), "dn_all": dn,
            "E": E.flatten(), "E2": E2.flatten()}
