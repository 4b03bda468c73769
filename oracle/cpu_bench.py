"""CPU-baseline timing of the reference path (test/bench infrastructure, NOT product code).

Executed only by bench.py's ``cpu_baseline`` leg and its ``--impl reference`` arm.  It times the
oracle port of the reference path (oracle/ref_path.py: sklearn BallTree + the per-point processors)
on a BOUNDED sample of query points against the FULL cloud's tree and extrapolates x N/Q, because the
full 10 M-point job takes hours on host cores.  The port hoists the scipy inverse-CDF
constants out of the per-point loop (the reference re-evaluates them per point, DM_saliency.py:99,
:282), so it is the reference's path at its most favourable.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import ref_path as rp

_G = {}


def _pass1_chunk(args):
    r, sub = args
    pts, tree, eps, min_points = _G["pts"], _G["tree"], _G["eps"], _G["min_points"]
    nb = tree.query_radius(pts[sub], r)
    out = np.zeros(len(sub))
    for t, (i, j) in enumerate(zip(sub, nb)):
        n, _ = rp.pca_point(pts[i], pts[j])
        k, _, _ = rp.curvature_point(pts[i], n, pts[j], eps, min_points)
        out[t] = k
    return out


def _pass2_chunk(args):
    r, sub = args
    pts, tree, nrm, K = _G["pts"], _G["tree"], _G["normals"], _G["K"]
    table, p = _G["table"], _G["params"]
    nb = tree.query_radius(pts[sub], r)
    out = np.zeros(len(sub))
    rho = r / np.sqrt(2.0)
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


class CpuSaliencyBaseline:
    """Tree built once over the full cloud; ``step`` times one bounded sample of the 3-radius job."""

    def __init__(self, pts, radii, params, workers=1, normals=None, K=None, seed=0):
        self.pts = np.ascontiguousarray(pts, np.float64)
        self.n = self.pts.shape[0]
        self.radii = tuple(radii)
        self.params = dict(params)
        self.workers = int(workers)
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
        _G.update(pts=self.pts, tree=self.tree, normals=self.normals, K=self.K,
                  eps=rp.curvature_epsilon(params["alpha"], params["roughness"]),
                  table=rp.chi2_table(params["alpha"], 8192), params=self.params,
                  min_points=params["min_points"])
        self.pool = None
        if self.workers > 1:
            self.pool = mp.get_context("fork").Pool(self.workers)   # tree/arrays shared copy-on-write

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def step(self, q_per_radius):
        """One bounded sample.  Returns (extrapolated seconds for the whole N-point 3-radius job
        including the one-off tree build, raw sample seconds)."""
        total = self.t_tree
        raw = 0.0
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
