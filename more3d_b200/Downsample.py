"""Saliency-driven and voxel-grid downsampling (the body of the reference's Downsampling/Downsample.py
script, :44-91, as functions) on the B200 kernels.

``saliency_downsample`` follows Downsample.py:44-83 with the DECLARED REPAIR of two lines that cannot run
as shipped: ``:58`` applies Python ``and`` to arrays (ValueError) and its two conjuncts
are mutually exclusive for any useful threshold -> ``idx_saliency = saliency > s_thresh``; ``:67`` indexes
with ``1 - bool_array`` (rows 0/1, not a complement) -> the mean runs over ``~idx_saliency``.
``literal=True`` keeps both conjuncts of :58 (``and`` -> ``&``) instead.
``voxel_downsample`` replaces ``open3d ... voxel_down_sample(voxel_size=LOD)`` (:86-89).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .BallTreePointSet import BallTreePointSet
from .grid import as_device_points


def threshold_otsu(values, nbins=256):
    """skimage.filters.threshold_otsu (0.19.x, README.md:90) of a device or host array: the 256-bin
    histogram is built on the GPU with numpy.histogram's exact binning; the 256-entry scan runs on the host."""
    lib = _lib.load()
    _lib.require_cuda()
    v = values if isinstance(values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(values))
    if v.dtype not in (torch.float32, torch.float64):
        v = v.double()
    v = v.to(device="cuda").contiguous().reshape(-1)
    if nbins != 256:
        raise NotImplementedError("only nbins=256 (the skimage default used by the reference)")
    n = int(v.numel())
    with torch.cuda.device(v.device):
        # extrema -> histogram over numpy.linspace(min, max, 257) entirely on the device, ONE host read at the end
        if v.dtype == torch.float32:
            mm = torch.empty(4, dtype=torch.float32, device=v.device)
            check(lib.more3d_minmax4(ptr(v), ptr(v), n, ptr(mm), stream_ptr()))
        else:
            mm = torch.empty(2, dtype=torch.float64, device=v.device)
            check(lib.more3d_cue_ops(4, ptr(v), n, None, 0.0, ptr(mm), stream_ptr()))
        hist = torch.empty(256, dtype=torch.int64, device=v.device)
        check(lib.more3d_histogram256_range(ptr(v), v.element_size(), n, ptr(mm), v.element_size(), ptr(hist), stream_ptr()))
        lohi = mm[:2].cpu().numpy().astype(np.float64)
    lo, hi = float(lohi[0]), float(lohi[1])
    if lo == hi:
        return lo
    edges = np.linspace(lo, hi, nbins + 1)          # numpy.histogram's bin edges for range=(min, max)
    counts = hist.cpu().numpy().astype(np.float64)
    centers = (edges[:-1] + edges[1:]) / 2.0
    with np.errstate(all="ignore"):
        w1 = np.cumsum(counts)
        w2 = np.cumsum(counts[::-1])[::-1]
        m1 = np.cumsum(counts * centers) / w1
        m2 = (np.cumsum((counts * centers)[::-1]) / w2[::-1])[::-1]
        var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return float(centers[int(np.argmax(var12))])


def voxel_downsample(pts, voxel_size, return_keys=False, return_device=False):
    """Mean point of every occupied voxel (Downsample.py:86-89).  Returns an (V,3) float64 array sorted by
    voxel index (open3d's own order is hash-map order: compare as sets), optionally the (V,3) int64 voxel
    indices."""
    lib = _lib.load()
    d = as_device_points(pts)
    n = int(d.shape[0])
    with torch.cuda.device(d.device):
        out = torch.empty((n, 3), dtype=torch.float64, device=d.device)
        keys = torch.empty((n, 3), dtype=torch.int64, device=d.device) if return_keys else None
        nout = C.c_int64(0)
        check(lib.more3d_voxel_downsample(ptr(d), n, float(voxel_size), ptr(out), ptr(keys), None, C.byref(nout),
                                          stream_ptr()))
    v = int(nout.value)
    out = out[:v]
    if not return_device:
        out = out.cpu().numpy()
    if return_keys:
        k = keys[:v] if return_device else keys[:v].cpu().numpy()
        return out, k
    return out


def saliency_downsample(pts, saliency, lod, leaf_size=40, zero_thresh=1e-4, literal=False, btP=None,
                        s_thresh=None, return_details=False, return_device=False):
    """One LOD of the loop at Downsample.py:44-83.  pts (N,3); saliency (N,).  Returns the (M,4) array the
    script writes (xyz + flag: 1 = salient / kept cell, 0 = representative of the averaged rest), rows in
    the reference's order (cell by cell, kept points in ball-tree order then the representative)."""
    lib = _lib.load()
    if btP is None:
        btP = BallTreePointSet(np.asarray(pts), leaf_size=leaf_size)
    tree = btP._hierarchy()
    dev = btP.device_points
    n = tree.n
    sal = saliency if isinstance(saliency, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(saliency))
    sal64 = sal.to(device=dev.device, dtype=torch.float64).contiguous().reshape(-1)
    if s_thresh is None:
        s_thresh = threshold_otsu(sal)                                      # :54
    cells_dev, cs = tree.smallest_nodes_dev(lod)                            # btP.getSmallestNodesOfSize(LOD), :47
    ncells = int(cells_dev.shape[0])
    with torch.cuda.device(dev.device):
        keep = torch.empty(n, dtype=torch.uint8, device=dev.device)
        keep_all = torch.empty(ncells, dtype=torch.uint8, device=dev.device)
        mean = torch.empty((ncells, 3), dtype=torch.float64, device=dev.device)
        check(lib.more3d_lod_cell_rule(ptr(dev), ptr(sal64), ptr(tree.idx_array_dev), ptr(cs), ncells, float(s_thresh),
                                       float(zero_thresh), 1 if literal else 0, ptr(keep), ptr(keep_all), ptr(mean),
                                       stream_ptr()))
        # closest original point to the mean of every averaged cell (btP.query(average_pt[None, :], 1), :70)
        avg = keep_all == 0
        near = None
        if bool(avg.any()):
            g = btP._grid(btP._knn_cell(1))
            near = g.knn(mean[avg].contiguous(), 1, return_distance=False)[0][:, 0].contiguous()
        cap = n + ncells
        out = torch.empty((cap, 4), dtype=torch.float64, device=dev.device)
        out_idx = torch.empty(cap, dtype=torch.int64, device=dev.device)
        out_cell = torch.empty(cap, dtype=torch.int32, device=dev.device)
        rows = C.c_int64(0)
        check(lib.more3d_lod_assemble(ptr(dev), ptr(tree.idx_array_dev), n, ptr(cs), ncells, ptr(keep), ptr(keep_all),
                                      ptr(near), ptr(out), ptr(out_idx), ptr(out_cell), C.byref(rows), stream_ptr()))
    m = int(rows.value)
    out, out_idx, out_cell = out[:m], out_idx[:m], out_cell[:m]
    if not return_device:
        out = out.cpu().numpy()
    if return_details:
        det = dict(idx=out_idx if return_device else out_idx.cpu().numpy(),
                   cell=out_cell if return_device else out_cell.cpu().numpy(), cells=cells_dev.cpu().numpy(),
                   s_thresh=s_thresh,
                   keep_all=keep_all.cpu().numpy().astype(bool))
        return out, det
    return out


if __name__ == '__main__':   # the reference script's driver (Downsample.py:30-91), paths as arguments
    import sys
    if len(sys.argv) < 3:
        raise SystemExit("usage: python -m more3d_b200.Downsample <saliency.txt (x y z s)> <out_prefix> [LOD ...]")
    saliency = np.loadtxt(sys.argv[1])
    pts = saliency[:, :3]
    lods = [float(a) for a in sys.argv[3:]] or [1, 2, 5, 7, 10]
    btP = BallTreePointSet(pts, leaf_size=40)
    for LOD in lods:
        print('computing {}'.format(LOD))
        np.savetxt(sys.argv[2] + '_' + str(LOD) + '_saliency.txt', saliency_downsample(pts, saliency[:, 3], LOD, btP=btP))
        voxel_down = voxel_downsample(pts, LOD)
        print('number of points', str(voxel_down.shape[0]))
        np.savetxt(sys.argv[2] + '_' + str(LOD) + '_voxel.txt', voxel_down)
