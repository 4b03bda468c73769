"""Sub-records of bench.py for BASELINE.json configs[0], [2], [3] (configs[1] is the bench line itself, configs[4]
its strong-scaling leg).  Each function runs the config at FULL size on the current GPU, times it with CUDA events
on the launching stream (best of a few repetitions after a warm-up), states the algorithmic bytes of SURVEY.md §8d
and the fraction of the measured HBM peak they amount to, times a BOUNDED sample of the reference's CPU path on the
host cores beside it, and spot-checks parity at that size.

    python tools/bench_configs.py [0] [2] [3] [--small]      # one JSON object per config

The CPU legs execute oracle/ (the checker) -- allowed for bench.py's CPU legs only; nothing here is product code.
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from more3d_b200 import synthetic, _lib  # noqa: E402


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def ev_ms(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), out


def best_ms(fn, reps=3):
    """best of `reps` timed calls after one warm-up (pool growth, module load)"""
    fn()
    best, out = 1e30, None
    for _ in range(reps):
        t, out = ev_ms(fn)
        best = min(best, t)
    return best, out


def frac(alg_bytes, ms, peak_gbs):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    return {"algorithmic_bytes": float(alg_bytes), "achieved_GBs": gbs, "frac_of_measured_hbm_peak": gbs / peak_gbs}


# ------------------------------------------------------------------------------------------------
def config0(n=1_000_000, peak=6545.9):
    """configs[0]: 1 M-point terrain, BallTreePointSet radius query r = 0.5 + PCA normals / curvature.
    Reference call sites: BallTreePointSet.py:35 (build), :317-347 (queryRadius)."""
    from more3d_b200.grid import UniformGrid, cell_for_radius
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from oracle import ref_path as rp
    pts = synthetic.terrain(n, seed=1)
    d = torch.from_numpy(pts).cuda()
    r = 0.5
    t_build, g = best_ms(lambda: UniformGrid(d, cell_for_radius(r)))
    t_q, (off, idx) = best_ms(lambda: g.radius_csr(None, r))
    t_n, (nrm, ratio, K, pc) = best_ms(lambda: g.normals_curvature(r, rp.curvature_epsilon(0.05, 0.02), 4))
    sum_m = float(off[-1].item())
    # the reference-shaped call: object array of int64 index arrays on the host (host-bound, reported as is)
    bt = BallTreePointSet(pts)
    bt.queryRadius(pts[:1000], r)
    t0 = time.perf_counter()
    api = bt.queryRadius(pts, r)
    t_api = time.perf_counter() - t0
    # CPU reference in full: sklearn BallTree (the call at BallTreePointSet.py:35) + query_radius (:345) over all
    # n queries, chunked over the host cores with a thread pool (query_radius releases the GIL; result identical)
    cores = host_cores()
    t0 = time.perf_counter()
    tree = rp.build_tree(pts)
    t_tree = time.perf_counter() - t0
    chunks = np.array_split(np.arange(n), max(cores * 8, 1))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        parts = list(ex.map(lambda c: tree.query_radius(pts[c], r), chunks))
    t_cpu_q = time.perf_counter() - t0
    want = np.concatenate(parts)
    cnt = np.fromiter((len(w) for w in want), np.int64, n)
    s1 = np.fromiter((w.sum() for w in want), np.int64, n)
    s2 = np.fromiter(((w.astype(np.float64) ** 2).sum() for w in want), np.float64, n)
    o = off.cpu().numpy()
    ix = idx.cpu().numpy().astype(np.int64)
    gc = np.diff(o)
    exact = bool(np.array_equal(gc, cnt) and np.array_equal(np.add.reduceat(ix, o[:-1]), s1)
                 and np.array_equal(np.add.reduceat(ix.astype(np.float64) ** 2, o[:-1]), s2))
    api_ok = bool(all(np.array_equal(np.sort(api[i]), np.sort(want[i])) for i in range(0, n, 997)))
    rec = {"workload": "configs[0]: %d-point terrain, BallTreePointSet radius query r=0.5 + PCA normals/curvature" % n,
           "points": n, "radius": r, "mean_m": sum_m / n,
           "gpu_ms": {"grid_build": t_build, "query_csr": t_q, "normals_curvature": t_n,
                      "total": t_build + t_q + t_n},
           "queries_per_s": n / (t_q * 1e-3),
           "query_csr": frac(n * 48.0 + 4.0 * sum_m, t_q, peak),
           "grid_build": frac(n * 176.0, t_build, peak),
           "normals_curvature": frac(n * 48.0 + 16.0 * sum_m, t_n, peak),
           "reference_api_s": {"BallTreePointSet.queryRadius(all points) -> object array on the host": t_api},
           "cpu": {"kind": "reference", "what": "sklearn BallTree build + query_radius over ALL %d queries "
                   "(thread pool over query chunks)" % n, "cores": cores, "tree_build_s": t_tree,
                   "query_radius_s": t_cpu_q, "queries_per_s": n / t_cpu_q},
           "parity": {"neighbour_sets_identical_all_rows(count, sum idx, sum idx^2)": exact,
                      "object_array_api_rows_identical(every 997th)": api_ok,
                      "pcount_equals_set_size": bool(np.array_equal(pc.cpu().numpy(), cnt))}}
    g.close()
    return rec


# ------------------------------------------------------------------------------------------------
def config2(n=50_000_000, peak=6545.9, cpu_points=1_000_000):
    """configs[2]: saliency-driven + voxel-grid downsampling at LOD 1, 2, 5, 7, 10 (Downsample.py:44-91)."""
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from more3d_b200.Downsample import saliency_downsample, voxel_downsample, threshold_otsu
    from more3d_b200.pipeline import multiscale_saliency
    from oracle import ref_path as rp
    d = synthetic.hash_terrain_device(3, 0, n, n)
    t_sal, res = ev_ms(lambda: multiscale_saliency(d, (0.5,)))
    sal_dev = res[0.5]
    del res
    pts = d.cpu().numpy()
    del d
    bt = BallTreePointSet(pts, leaf_size=40)
    torch.cuda.synchronize()
    t_tree, tree = ev_ms(lambda: bt._hierarchy())
    t_otsu, thr = ev_ms(lambda: threshold_otsu(sal_dev))
    levels = int(np.log2(tree.n_nodes + 1))
    rec = {"workload": "configs[2]: saliency-driven + voxel-grid downsampling of %d points at LOD 1/2/5/7/10" % n,
           "points": n, "saliency_input": "DM_saliency pipeline at r=0.5 on the same cloud (%.1f ms)" % t_sal,
           "tree_build_ms": t_tree, "n_nodes": int(tree.n_nodes),
           "tree_build": frac(n * 60.0 * levels, t_tree, peak), "otsu_ms": t_otsu, "otsu_threshold": thr, "lods": {}}
    tot = t_tree + t_otsu
    for lod in (1, 2, 5, 7, 10):
        t_dev, out = best_ms(lambda: saliency_downsample(pts, sal_dev, lod, btP=bt, s_thresh=thr, return_device=True), reps=2)
        kept = int(out.shape[0])
        del out
        t_vox, vox = best_ms(lambda: voxel_downsample(bt.device_points, float(lod), return_device=True), reps=2)
        nv = int(vox.shape[0])
        del vox
        tot += t_dev + t_vox
        rec["lods"][str(lod)] = {"saliency_downsample_ms": t_dev, "kept_rows": kept, "voxel_ms": t_vox, "voxels": nv,
                                 "voxel": frac(n * (24.0 + 12.0 * 9 + 24.0) + nv * 24.0, t_vox, peak)}
    rec["total_ms"] = tot
    rec["points_per_s"] = n / (tot * 1e-3)
    # parity spot-check at size: voxel means of LOD 5 vs the oracle on the voxels of a 200 k-point window
    keys_ok = None
    try:
        out, keys = voxel_downsample(bt.device_points, 5.0, return_keys=True, return_device=True)
        vmin = pts.min(axis=0) - 2.5
        kk = np.floor((pts - vmin) / 5.0).astype(np.int64)
        sel = np.flatnonzero((kk[:, 0] == kk[0, 0]) & (kk[:, 1] == kk[0, 1]))
        vk = kk[0]
        row = torch.nonzero((keys[:, 0] == int(vk[0])) & (keys[:, 1] == int(vk[1]))).flatten().cpu().numpy()
        got = out[row].cpu().numpy()
        gk = keys[row].cpu().numpy()
        okk = True
        for z in np.unique(kk[sel, 2]):
            m = sel[kk[sel, 2] == z]
            s = np.zeros(3)
            for i in m:              # sequential fp64 sum in point order, as the oracle's np.add.at
                s = s + pts[i]
            j = np.flatnonzero(gk[:, 2] == z)
            okk = okk and len(j) == 1 and np.array_equal(got[j[0]], s / len(m))
        keys_ok = bool(okk)
        del out, keys
    except Exception as e:  # pragma: no cover
        keys_ok = "not checked: %s" % e
    rec["parity"] = {"voxel_column_means_bit_equal_to_sequential_fp64_sum": keys_ok}
    # bounded CPU leg: the oracle (sklearn tree + per-cell loop, open3d-style voxel binning) on a sub-tile
    m = min(cpu_points, n)
    sub = np.ascontiguousarray(pts[:m])
    ssub = sal_dev[:m].cpu().numpy().astype(np.float64)
    t0 = time.perf_counter()
    ctree = rp.build_tree(sub, 40)
    t_ctree = time.perf_counter() - t0
    t0 = time.perf_counter()
    rp.saliency_downsample(sub, ssub, 5.0, tree=ctree)
    t_csal = time.perf_counter() - t0
    t0 = time.perf_counter()
    rp.voxel_downsample(sub, 5.0)
    t_cvox = time.perf_counter() - t0
    rec["cpu"] = {"kind": "port", "what": "oracle restatement (sklearn BallTree build + Downsample.py loop, declared repair; "
                  "open3d-style voxel binning) on the first %d points, ONE LOD (5)" % m, "cores": 1, "points": m,
                  "tree_build_s": t_ctree, "saliency_downsample_s": t_csal, "voxel_s": t_cvox,
                  "points_per_s_one_lod": m / (t_csal + t_cvox)}
    del bt, tree, sal_dev
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------
def config3(n=20_000_000, iters=200, peak=6545.9, cpu_points=1_000_000):
    """configs[3]: point-based level set, k = 7, h = 0.5, Chan-Vese, 200 iterations = 4 x run(50)
    (levelsets_func.py:598-642; run_levelsets.py:68-69, 240-245)."""
    from more3d_b200 import levelsets_func as ls
    from more3d_b200.pipeline import multiscale_saliency
    d = synthetic.hash_terrain_device(4, 0, n, n)
    res = multiscale_saliency(d, (0.5,), keep=True)
    zeta = res[0.5]["_nonparam_K"].abs().double().cpu().numpy()[:, None]
    del res
    pts = d.cpu().numpy()
    del d
    pts_c = pts - pts.mean(axis=0)
    h, k = 0.5, 7
    t0 = time.perf_counter()
    nb, nrm, tang = ls.build_neighborhoods(pts_c, h, k, return_device=True)
    torch.cuda.synchronize()
    t_nb = time.perf_counter() - t0
    t0 = time.perf_counter()
    solver = ls.Solver(h, zeta, pts_c, nb, tang, "chan_vese")
    torch.cuda.synchronize()
    t_solver = time.perf_counter() - t0
    ls.normalize(solver.zeta)
    solver.initialize_from_voxels(2.0)
    a0 = int(solver.active.sum())
    run = dict(stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
    per = max(iters // 4, 1)
    times, rms = [], []
    for _ in range(4):
        t, _ = ev_ms(lambda: solver.run(per, **run))
        times.append(t)
        rms.append(solver.rmsc_history[-1])
    phi = solver.phi
    a1 = int(solver.active.sum())
    ms_it = sum(times) / (4 * per)
    n_act = 0.5 * (a0 + a1)
    rec = {"workload": "configs[3]: point-based level set (Chan-Vese, curvature cue) on %d points, k=7, %d iterations" % (n, 4 * per),
           "points": n, "k": k, "h": h, "neighbourhoods_s": t_nb, "solver_init_s": t_solver,
           "singular_operators": int(solver.diff.singular), "run_ms_per_cycle": times, "iterations": 4 * per,
           "ms_per_iteration": ms_it, "point_iterations_per_s": n / (ms_it * 1e-3),
           "active_start": a0, "active_end": a1, "rmsc_last_per_cycle": rms,
           "iteration": frac(n_act * 1800.0 + n * 101.0, ms_it, peak),
           "properties": {"phi_within_cap": bool(np.abs(phi).max() <= 2 * h + 1e-12), "phi_finite": bool(np.all(np.isfinite(phi)))}}
    del solver, nb, nrm, tang
    torch.cuda.empty_cache()
    # bounded CPU leg: the level-set oracle (numpy restatement of levelsets_func) on a sub-tile, 3 iterations
    from oracle import ref_levelset as rl
    m = min(n, cpu_points)
    sub = synthetic.hash_terrain_host(4, 0, m, m)
    sub -= sub.mean(axis=0)
    t0 = time.perf_counter()
    cnb, cn, ct1, ct2 = rl.knn_frames(sub, k)
    t_cpu_nb = time.perf_counter() - t0
    t0 = time.perf_counter()
    cw, cD = rl.wls_operator(sub, cnb, ct1, ct2, 3 * h)
    t_cpu_op = time.perf_counter() - t0
    cphi = rl.init_voxels(sub, 2.0, 2 * h)
    cact = rl.zero_crossings(cphi, cnb, np.arange(m))
    st = dict(points=sub, nbrs=cnb, t1=ct1, t2=ct2, w=cw, D=cD, h=h, zeta=np.abs(np.sin(sub[:, :1])), phi=cphi, active=cact,
              last_active=cact, alias=True)
    t0 = time.perf_counter()
    rl.run(st, 3, **run)
    t_cpu_it = (time.perf_counter() - t0) / 3
    rec["cpu"] = {"kind": "port", "what": "oracle restatement of levelsets_func (numpy/LAPACK) on %d points, 3 iterations" % m,
                  "cores": "numpy/LAPACK threads", "points": m, "neighbourhoods_s": t_cpu_nb, "operator_s": t_cpu_op,
                  "s_per_iteration": t_cpu_it, "point_iterations_per_s": m / t_cpu_it}
    return rec


def run_all(which=("0", "2", "3"), small=False, peak=6545.9):
    out = {}
    if "0" in which:
        out["0"] = config0(200_000 if small else 1_000_000, peak=peak)
    if "2" in which:
        out["2"] = config2(int(float(os.environ.get("MORE3D_BENCH_N2", 1e6 if small else 5e7))), peak=peak,
                           cpu_points=200_000 if small else 1_000_000)
    if "3" in which:
        out["3"] = config3(500_000 if small else 20_000_000, 40 if small else 200, peak=peak,
                           cpu_points=50_000 if small else 250_000)
    return out


if __name__ == "__main__":
    small = "--small" in sys.argv
    which = [a for a in sys.argv[1:] if a.isdigit()] or ["0", "2", "3"]
    _lib.load()
    for k, v in run_all(which, small).items():
        print(json.dumps({"config": k, **v}), flush=True)
