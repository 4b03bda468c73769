"""The evaluation of the reference's ``Downsampling/downsample_compare.py`` (:51-64, :102-150) as functions
on the B200 kernels: PCA normals by radius or kNN with the upward-z flip, symmetric 1-NN distances,
cloud-to-cloud / cloud-to-plane errors and the median normal deviation.  ("Next" row f-2 of SURVEY.md §8;
the script itself loads four files written by other software and is out of scope.)

open3d's ``estimate_normals`` is absent offline: normals here are the centroid-PCA normals of the same
neighbourhoods (parity unpinned), ``(0, 0, 1)`` where fewer than 3 neighbours exist (open3d's default).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .grid import UniformGrid, X_SPLIT, as_device_points, cell_for_radius


def _knn_cell(d, k):
    b = _lib.cloud_bounds(d)
    ext = b[3:] - b[:3]
    cell = float(np.sqrt(max(k, 8) * max(float(ext[0]) * float(ext[1]), 1e-300) / (4.0 * d.shape[0])))
    return cell if np.isfinite(cell) and cell > 0 else 1.0


def estimate_normals(points, radius=None, knn=None, upwards_z=True, return_device=False):
    """``estimate_normals(KDTreeSearchParamRadius(radius))`` or ``...KNN(knn)`` + the upward-z flip
    (downsample_compare.py:53-62)."""
    lib = _lib.load()
    d = as_device_points(points)
    n = int(d.shape[0])
    with torch.cuda.device(d.device):
        if radius is not None:
            g = UniformGrid(d, cell_for_radius(radius), x_split=X_SPLIT)
            normals, _, _ = g.normals_pca(radius)
            g.close()
        else:
            k = int(min(knn, n))
            g = UniformGrid(d, _knn_cell(d, k))
            idx, _ = g.knn(None, k, return_distance=False)
            g.close()
            normals = torch.empty((n, 3), dtype=torch.float64, device=d.device)
            check(lib.more3d_normals_from_knn(ptr(d), n, k, ptr(idx), ptr(normals), None, stream_ptr()))
        check(lib.more3d_orient_normals(ptr(normals), n, 1 if upwards_z else 0, stream_ptr()))
    return normals if return_device else normals.cpu().numpy()


def _nearest(src, dst):
    """for every row of dst: distance to and index of its nearest neighbour in src (BallTree.query(dst, 1))"""
    g = UniformGrid(src, _knn_cell(src, 8))
    idx, dist = g.knn(dst, 1, return_distance=True)
    g.close()
    return dist[:, 0], idx[:, 0].contiguous()


def compare(pts, downsampled, normals_pts=None, rnn=.1, knn=64, rnn_flag=True, upwards_z=True):
    """One (original, downsampled) pair of the loop at downsample_compare.py:102-150.
    Returns dict(c2c, c2p, dN, dn_all, E, E2) with the reference's expressions kept as written
    (``np.average(E2) + np.average(E) / 2`` at :142-143; the normals indexing of :145)."""
    lib = _lib.load()
    P = as_device_points(np.asarray(pts)[:, :3] if not isinstance(pts, torch.Tensor) else pts[:, :3])
    Q = as_device_points(np.asarray(downsampled)[:, :3] if not isinstance(downsampled, torch.Tensor) else downsampled[:, :3])
    if normals_pts is None:
        normals_pts = estimate_normals(P, radius=rnn if rnn_flag else None, knn=knn, upwards_z=upwards_z, return_device=True)
    NP = as_device_points(normals_pts)
    NQ = estimate_normals(Q, radius=rnn if rnn_flag else None, knn=knn, upwards_z=upwards_z, return_device=True)
    with torch.cuda.device(P.device):
        check(lib.more3d_orient_normals(ptr(NQ), int(NQ.shape[0]), 2, stream_ptr()))   # normals.normalize_normals(), :122
        E, closest = _nearest(P, Q)                              # btP.query(downsampled, 1), :132
        E2, closest2 = _nearest(Q, P)                            # downsampled_bt.query(pts, 1), :136
        c2p = torch.empty(P.shape[0], dtype=torch.float64, device=P.device)
        ang = torch.empty(Q.shape[0], dtype=torch.float64, device=P.device)
        check(lib.more3d_compare_terms(ptr(P), int(P.shape[0]), ptr(Q), int(Q.shape[0]), ptr(closest), ptr(closest2),
                                       ptr(NP), ptr(NQ), ptr(c2p), ptr(ang), stream_ptr()))
        # np.median: mean of the two middle ORDER STATISTICS, found by radix select on the device (no sort)
        m = int(ang.numel())
        mid = torch.empty(2, dtype=torch.float64, device=P.device)
        ranks = (C.c_int64 * 2)((m - 1) // 2, m // 2)
        check(lib.more3d_select_kth(ptr(ang), m, ranks, 2, ptr(mid), stream_ptr()))
        # np.average of the three error columns: library reductions (fixed tree), read back together with the median
        avg = torch.empty(3, dtype=torch.float64, device=P.device)
        for i, col in enumerate((E, E2, c2p)):
            col = col.reshape(-1).contiguous()
            check(lib.more3d_cue_ops(5, ptr(col), int(col.numel()), None, 0.0, ptr(avg[i:i + 1]), stream_ptr()))
    mid = mid.cpu().numpy()
    avg = avg.cpu().numpy()
    med = (float(mid[0]) + float(mid[1])) / 2.0
    return {"c2c": float(avg[1] + avg[0] / 2), "c2p": float(avg[2]), "dN": float(med),
            "dn_all": ang.cpu().numpy(), "E": E.cpu().numpy(), "E2": E2.cpu().numpy()}
