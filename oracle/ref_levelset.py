"""CPU ORACLE for the point-based level set -- test / bench infrastructure, NOT product code.

A functional numpy restatement of the numerical part of the reference's ``levelsets_func.py``
(frames :112-151, WLS operator :246-275, derivative application :277-305, smooth :350-358,
Chan-Vese step :386-437, local-mean-and-variance step :439-513, run loop :598-642).  Pinned by
``tests/golden/levelset_4k.npz``, ``tests/golden/levelset_mask_3k.npz`` and ``tests/golden/levelset_lmv_3k.npz``
(the reference executed verbatim, ``tools/make_golden.py``) in
``tests/test_oracle_golden.py``.  Only tests and ``tools/bench_configs.py``'s CPU-timing leg import it.
"""
import numpy as np
from sklearn.neighbors import BallTree


def knn_frames(points, k, normals=None):
    """build_neighborhoods (:154-172): 2k-NN (self first), PCA normal of the offsets about the query point,
    tangent frame from the farthest of the 2k neighbours.  Returns (nbrs (N,k), normals, t1, t2)."""
    tree = BallTree(points, leaf_size=64)
    nn = tree.query(points, 2 * k, return_distance=False, sort_results=True, dualtree=True)
    if normals is None:
        off = points[nn] - points[:, None, :]
        normals = np.linalg.svd(off)[2][:, -1]
    n = normals / np.linalg.norm(normals, axis=1)[:, None]
    t1 = points - points[nn[:, -1]]
    t1 = t1 - np.sum(n * t1, axis=1)[:, None] * n
    t1 /= np.linalg.norm(t1, axis=1)[:, None]
    t2 = np.cross(n, t1)
    t2 /= np.linalg.norm(t2, axis=1)[:, None]
    return nn[:, 1:k + 1], n, t1, t2


def wls_operator(points, nbrs, t1, t2, max_d):
    """Derivative.__init__ (:246-275): weights (N,k) and D = pinv(M) (N,5,k+3)."""
    n, k = nbrs.shape
    d = points[nbrs] - points[:, None, :]
    x = np.einsum("nkc,nc->nk", d, t1)
    y = np.einsum("nkc,nc->nk", d, t2)
    dist = np.linalg.norm(d, axis=-1)
    wend = np.where(dist < max_d, (1 - dist / max_d) ** 4 * (4 * dist / max_d + 1), 0.0)
    w = (np.sqrt(wend) + 1e-8) / (1 + 1e-8)
    M = np.zeros((n, k + 3, 5))
    M[:, :k] = np.stack([x, y, x * y, x * x, y * y], axis=-1) * w[:, :, None]
    M[:, k, 2] = M[:, k + 1, 3] = M[:, k + 2, 4] = 1.0
    U, S, Vh = np.linalg.svd(M, full_matrices=False)
    D = np.einsum("nij,nki->njk", Vh, U / S[:, None, :])
    return w, D


def apply(D, w, nbrs, u, rows, idx):
    """Derivative.__call__ (:277-291) restricted to operator rows `rows` at the points `idx`."""
    f = np.empty((len(idx), nbrs.shape[1] + 3))
    f[:, :-3] = w[idx] * (u[nbrs[idx]] - u[idx, None])
    f[:, -3:] = 1.0 / 1000
    return np.einsum("nrj,nj->nr", D[idx][:, rows, :], f)


def smooth(u, idx, nbrs, w):
    """:350-358"""
    return (u[nbrs[idx]] * w[idx]).sum(axis=1) / w[idx].sum(axis=1)


def heaviside(x, eps):
    return 0.5 * (1 + 2 / np.pi * np.arctan(x / eps))


def double_well(s):
    out = np.empty_like(s)
    lo = s < 1
    out[lo] = np.sinc(2 * s[lo])
    out[~lo] = 1 - 1 / s[~lo]
    return out


def init_voxels(points, h, max_val):
    """initialize_from_voxels (:591-594)."""
    even = np.sum(points // h, axis=-1) % 2 == 0
    return np.where(even, -max_val, max_val)


def zero_crossings(phi, nbrs, idx):
    return np.any(np.sign(phi[idx])[:, None] != np.sign(phi[nbrs[idx]]), axis=1)


def lmv_force(phi, zeta, nb, idx, eps):
    """The data term of _lmv_step (:455-512) at the points idx; zeta is the 1-D cue the step is written for."""
    n = len(phi)
    H = heaviside(phi, eps)
    hood = np.column_stack((np.arange(n), nb))                      # self first, :383-384
    m = hood.shape[1]
    sel = np.zeros(n, bool)
    sel[idx] = True
    sel[hood[idx]] = True                                           # idx_, :463-464
    hs = hood[sel]
    mu_in, mu_out, s_in, s_out = (np.zeros(n) for _ in range(4))
    mu_in[sel] = (zeta[hs] * (1 - H[hs])).sum(axis=-1) / (1 - H[hs]).sum(axis=-1)
    mu_out[sel] = (zeta[hs] * H[hs]).sum(axis=-1) / H[hs].sum(axis=-1)
    s_in[sel] = (((zeta[hs] - mu_in[sel, None]) ** 2 * (1 - H[hs])).sum(axis=-1) / (1 - H[hs]).sum(axis=-1)) ** 0.5
    s_out[sel] = (((zeta[hs] - mu_out[sel, None]) ** 2 * H[hs]).sum(axis=-1) / H[hs].sum(axis=-1)) ** 0.5
    s_in[sel] = np.maximum(s_in[sel], 0.0001)
    s_out[sel] = np.maximum(s_out[sel], 0.0001)
    hi = hood[idx]
    e_in = 1 / m * (np.log(np.sqrt(2 * np.pi) * s_out[idx, None])
                    + (zeta[hi] - mu_out[hi]) ** 2 / (2 * s_out[idx, None] ** 2)).sum(axis=-1)
    e_out = 1 / m * (np.log(np.sqrt(2 * np.pi) * s_in[idx, None])
                     + (zeta[hi] - mu_in[hi]) ** 2 / (2 * s_in[idx, None] ** 2)).sum(axis=-1)
    return e_in - e_out


def run(state, iterations, stepsize, nu, mu, lambda1, lambda2, epsilon, cap=True, tolerance=1e-4, model="chan_vese"):
    """Solver.run (:598-642) on a state dict with keys points, nbrs, t1, t2, w, D, h, zeta (N,C), phi, active,
    last_active, alias (True when last_active IS active, :581/:589/:596).  Mutates phi / active / last_active.
    model='lmv': _lmv_step (:439-513); zeta is then 1-D (or (N,1)) and lambda1 / lambda2 are unused."""
    lmv = model == "lmv"
    if lmv:
        state["zeta"] = np.asarray(state["zeta"], dtype=np.float64).reshape(len(state["phi"]))
    nb, w, D, t1, t2 = state["nbrs"], state["w"], state["D"], state["t1"], state["t2"]
    phi, zeta, h = state["phi"], state["zeta"], state["h"]
    max_val = 2 * h
    eps = epsilon * h
    n = len(phi)
    for _ in range(iterations):
        active = state["active"]
        if not state["alias"]:
            changed = active != state["last_active"]
            phi[changed] = np.sign(phi[changed]) * max_val          # :603-608 + initialize_new
        idx = np.nonzero(active)[0]
        R = C = 0.0
        if nu > 0 or mu > 0 or lmv:                                 # _lmv_step differentiates unconditionally
            a = apply(D, w, nb, phi, [0, 1, 3, 4], idx)
            lap = 2 * a[:, 2] + 2 * a[:, 3]
            grad = a[:, :1] * t1[idx] + a[:, 1:2] * t2[idx]
            mag = np.linalg.norm(grad, axis=1)
            keep = mag > 0
            idx, lap, grad, mag = idx[keep], lap[keep], grad[keep], mag[keep]
            dpf = np.zeros(n)
            dpf[idx] = double_well(mag)
        if nu > 0 or lmv:
            g = apply(D, w, nb, dpf, [0, 1], idx)
            gd = g[:, :1] * t1[idx] + g[:, 1:2] * t2[idx]
            R = nu * (dpf[idx] * lap + np.sum(gd * grad, axis=1))
        if not lmv:
            H = heaviside(phi, eps)[:, None]
            z_in = (zeta * (1 - H)).sum(axis=0) / (1 - H).sum()
            z_out = (zeta * H).sum(axis=0) / H.sum()
        if mu > 0 or lmv:
            unit = np.zeros((n, 3))
            unit[idx] = grad / mag[:, None]
            div = 0.0
            for c in range(3):
                g = apply(D, w, nb, unit[:, c], [0, 1], idx)
                div = div + (g[:, 0] * t1[idx, c] + g[:, 1] * t2[idx, c])
            C = mu * div
        dlt = 1 / np.pi * eps / (eps ** 2 + phi[idx] ** 2)
        if lmv:
            F = lmv_force(phi, zeta, nb, idx, eps)
        else:
            F = lambda1 * np.sum((zeta[idx] - z_in) ** 2, axis=1) - lambda2 * np.sum((zeta[idx] - z_out) ** 2, axis=1)
        inc = stepsize * (R + dlt * (C + F))
        phi[idx] += inc
        rmsc = np.mean(inc ** 2) ** 0.5
        if not state["alias"]:
            state["last_active"] = active.copy()
        act_idx = np.nonzero(active)[0]
        new_active = np.zeros(n, bool)
        new_active[act_idx] = zero_crossings(phi, nb, act_idx)
        if (mu > 0 or nu > 0) and cap:
            out = np.nonzero((phi > max_val) | (phi < -max_val))[0]
            phi[out] = np.sign(phi[out]) * max_val
            phi[out] = smooth(phi, out, nb, w)
        lonely = np.nonzero(np.all(np.sign(phi[nb]) != np.sign(phi)[:, None], axis=1))[0]
        phi[lonely] = smooth(phi, lonely, nb, w)
        new_active[nb[new_active]] = True
        state["active"] = new_active
        if state["alias"]:
            state["last_active"] = new_active
        if rmsc < tolerance:
            return True
    return False
