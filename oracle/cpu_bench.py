"""CPU-baseline timing of the reference path (test/bench infrastructure, NOT product code).

Executed only by bench.py's ``cpu_baseline`` leg and its ``--impl reference`` arm.  Two kinds:

``kind="reference"`` -- the UNMODIFIED reference (oracle/_ref bytecode, oracle/ref_verbatim.py): the tree call the
    reference makes (``sklearn.neighbors.BallTree(points, leaf_size=40)`` + ``query_radius``,
    BallTreePointSet.py:35, :339-345) and the verbatim ``Curvature_Kernel.process`` / ``Saliency_Kernel.process``
    callbacks (DM_saliency.py:88-157, :219-324), scipy inverse CDFs evaluated per point as the reference does.
    ``BallTreePointSet.__init__`` itself is not used at bench sizes: its O(nodes^2) Python hierarchy
    reconstruction (:45-53) takes hours at 10 M points and is not part of the saliency path.
``kind="port"`` -- oracle/ref_path.py, the numpy restatement (scipy constants hoisted: the reference's path at
    its most favourable); the fallback when oracle/_ref is absent.

Both run on a BOUNDED uniform random sample of Q query points per radius against the FULL cloud's tree and
extrapolate x N/Q (the full 10 M-point job takes hours on host cores).  The PCA
normals the reference takes as a precondition (OPALS Normals module, DM_saliency.py:329-330) are computed -- and
timed -- with the oracle's centroid-PCA for the sampled points.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import ref_path as rp
from . import ref_verbatim as rv

_G = {}


def _pass1_chunk(args):
    r, sub = args
    pts, tree, p = _G["pts"], _G["tree"], _G["params"]
    nb = tree.query_radius(pts[sub], r)                       # BallTreePointSet.py:345
    nrm = np.empty((len(sub), 3))
    for t, (i, j) in enumerate(zip(sub, nb)):
        nrm[t], _ = rp.pca_point(pts[i], pts[j])
    if _G["kind"] == "reference":
        K, _, _ = rv.run_curvature(pts, nrm, nb, p["alpha"], p["roughness"], p["min_points"], queries=sub)
        return K
    out = np.zeros(len(sub))
    for t, (i, j) in enumerate(zip(sub, nb)):
        out[t], _, _ = rp.curvature_point(pts[i], nrm[t], pts[j], _G["eps"], p["min_points"])
    return out


def _pass2_chunk(args):
    r, sub = args
    pts, tree, nrm, K = _G["pts"], _G["tree"], _G["normals"], _G["K"]
    table, p = _G["table"], _G["params"]
    nb = tree.query_radius(pts[sub], r)
    rho = r / np.sqrt(2.0)
    if _G["kind"] == "reference":
        dk, dn = rv.run_dk_dn(pts, nrm, K, nb, rho, rho, p["sigma_normal"], p["sigma_curvature"], p["alpha"],
                              p["min_points"], queries=sub)
        return dk + dn
    out = np.zeros(len(sub))
    for t, (i, j) in enumerate(zip(sub, nb)):
        dk, dn = rp.saliency_point(pts[i], nrm[i], K[i], pts[j], nrm[j], K[j], rho, p["sigma_normal"],
                                   p["sigma_curvature"], table, p["min_points"])
        out[t] = dk + dn
    return out


def _run(fn, r, sub, workers, pool):
    chunks = [(r, c) for c in np.array_split(sub, max(workers * 4, 1)) if len(c)]
    t0 = time.perf_counter()
    if pool is None:
        for c in chunks:
            fn(c)
    else:
        pool.map(fn, chunks)
    return time.perf_counter() - t0


def best_kind():
    return "reference" if rv.available() else "port"


def describe(kind):
    if kind == "reference":
        return ("UNMODIFIED reference (%s): sklearn BallTree(leaf_size=40).query_radius as called at "
                "BallTreePointSet.py:35,:345 + verbatim Curvature_Kernel.process / Saliency_Kernel.process "
                "(DM_saliency.py:88-157, :219-324) per point; precondition normals by centroid-PCA" % rv.origin())
    return ("oracle port (sklearn BallTree.query_radius + per-point PCA/Curvature_Kernel/Saliency_Kernel numpy "
            "restatement, scipy constants hoisted)")


class CpuSaliencyBaseline:
    """Tree built once over the full cloud; ``step`` times one bounded sample of the multi-radius job."""

    def __init__(self, pts, radii, params, workers=1, normals=None, K=None, seed=0, kind=None):
        self.pts = np.ascontiguousarray(pts, np.float64)
        self.n = self.pts.shape[0]
        self.radii = tuple(radii)
        self.params = dict(params)
        self.workers = int(workers)
        self.kind = kind or best_kind()
        if self.kind == "reference":
            rv.load_reference()                            # import cost outside the timed region
        t0 = time.perf_counter()
        self.tree = rp.build_tree(self.pts, 40)           # BallTreePointSet.py:35
        self.t_tree = time.perf_counter() - t0
        rng = np.random.Generator(np.random.PCG64(seed))
        if normals is None:
            # stand-in attributes for the dk/dn timing leg: upward normals with a small tilt and
            # ~10 % non-zero curvature values (cost is value-independent to first order)
            normals = np.zeros((self.n, 3))
            normals[:, :2] = rng.normal(0, 0.05, (self.n, 2))
            normals[:, 2] = 1.0
            K = np.where(rng.random(self.n) < 0.1, rng.normal(0, 0.05, self.n), 0.0)
        self.normals = np.ascontiguousarray(normals, np.float64)
        self.K = np.asarray(K, np.float32).astype(np.float64)
        self.rng = rng
        _G.update(pts=self.pts, tree=self.tree, normals=self.normals, K=self.K, kind=self.kind,
                  eps=rp.curvature_epsilon(params["alpha"], params["roughness"]),
                  table=rp.chi2_table(params["alpha"], 8192), params=self.params)
        self.pool = None
        if self.workers > 1:
            self.pool = mp.get_context("fork").Pool(self.workers)   # tree/arrays shared copy-on-write

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def step(self, q_per_radius):
        """One bounded sample.  Returns (extrapolated seconds for the whole N-point multi-radius job
        including the one-off tree build, raw sample seconds)."""
        total = self.t_tree
        raw = 0.0
        q_per_radius = int(min(q_per_radius, self.n))
        for r in self.radii:
            sub = self.rng.choice(self.n, q_per_radius, replace=False)
            t1 = _run(_pass1_chunk, r, sub, self.workers, self.pool)
            t2 = _run(_pass2_chunk, r, sub, self.workers, self.pool)
            raw += t1 + t2
            total += (t1 + t2) * (self.n / float(q_per_radius))
        return total, raw


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
