"""Device-resident uniform grid: thin host wrapper over the C-ABI (include/more3d_b200.h).

torch is used for device memory and streams only; all compute is in libmore3d_b200.so.
"""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

import os

CELL_MARGIN = 1.0 + 1.0 / 65536.0
# Grid columns per query radius.  1 -> 3x3 column stencil (candidate/neighbour ratio 9/pi = 2.9);
# 2 -> 5x5 columns of half the width (25/(4 pi) = 2.0): fewer distance tests per neighbour at the price of
# more, shorter row ranges.  Measured on B200, 10 M-point terrain: 2 is 17 % faster than 1 over the
# moments + dk/dn kernels, 3 is no better than 2 (profiles/README.md).
CELLS_PER_RADIUS = int(os.environ.get("MORE3D_CELLS_PER_RADIUS", "2"))
# Columns of the fixed-radius grids are this many times narrower along x than they are high along y (see
# more3d_grid_create_segments): configs[1] takes 22.6 / 21.9 / 21.3 ms per step with 1 / 2 / 4 (the column table grows
# by the same factor: 4 columns per point at 4).
X_SPLIT = int(os.environ.get("MORE3D_X_SPLIT", "4"))


def cell_for_radius(radius):
    """Column width used for fixed-radius work at `radius`."""
    return max(float(radius), 1e-300) * CELL_MARGIN / CELLS_PER_RADIUS


def as_device_points(points, device=None):
    """(N,3) float64 C-order CUDA tensor from numpy / torch input (host inputs are copied H2D)."""
    _lib.require_cuda()
    if isinstance(points, torch.Tensor):
        t = points
    else:
        a = np.ascontiguousarray(points, dtype=np.float64)
        t = torch.from_numpy(a)
    if t.ndim != 2 or t.shape[1] != 3:
        raise ValueError("points must have shape (N, 3)")
    dev = torch.device(device if device is not None else "cuda")
    return t.to(device=dev, dtype=torch.float64).contiguous()


class GridColumn:
    """A per-point column stored in the sorted order of a grid.  `tensor()` returns it in the caller's point
    order (one un-permute pass, cached).  The neighbourhood passes write their outputs this way because sorted
    writes are coalesced; consumers that are order-agnostic (the next pass on the same grid, the blend, min/max)
    read `.data` directly and the un-permutation never happens for intermediate columns nobody looks at."""

    def __init__(self, grid, data):
        # weak: the grid remembers its latest columns (_feat_cols); a strong back-reference would make a cycle
        # and keep hundreds of MB of HBM alive until the cyclic collector happens to run
        if isinstance(grid, GridColumn):     # another column of the same grid
            self._grid, self.order = grid._grid, grid.order
        else:
            self._grid = weakref.ref(grid)
            self.order = grid.order_tensor()     # shared by all columns of the grid; outlives grid.close()
        self.data = data
        self._orig = None
        self._v0 = data._version

    @property
    def grid(self):
        return self._grid()

    def tensor(self):
        if self._orig is None:
            t = self.data.contiguous()
            with torch.cuda.device(t.device):
                out = torch.empty_like(t)
                nbytes = t.element_size() * (int(t.shape[1]) if t.ndim == 2 else 1)
                check(_lib.load().more3d_unpermute(ptr(self.order), int(t.shape[0]), ptr(t), ptr(out), int(nbytes),
                                                   stream_ptr()))
            self._orig = out
            self._v1 = out._version
        return self._orig

    def clean(self):
        """True while neither representation has been modified in place since it was produced."""
        return self.data._version == self._v0 and (self._orig is None or self._orig._version == self._v1)


class UniformGrid:
    """Spatial index replacing the reference's ball tree (BallTreePointSet.py:35)."""

    def __init__(self, xyz, cell, left=None, right=None, lattice=None, bounds=None, x_split=1):
        """xyz: (N,3) cloud.  Strips (multi-GPU): `left` / `right` are the ghost points received from the two
        neighbours -- the grid indexes the three tensors in place as one cloud [left, xyz, right] (no concatenated
        copy); `lattice` = (origin_x, origin_y) of the column lattice shared by all strips of the tile (bit-identical
        binning, ordering and sums to the whole-tile grid); `bounds` = (min x, y, z, max x, y, z) if already known;
        `x_split`: columns cell / x_split wide (fewer rejected candidates per stencil row; kNN needs 1)."""
        self.lib = _lib.load()
        self.xyz = as_device_points(xyz)
        self.device = self.xyz.device
        segs = [t for t in (left, self.xyz, right) if t is not None and int(t.shape[0]) > 0]
        segs = [as_device_points(t, self.device) for t in segs]
        self._segs = segs                                   # the grid reads them during the build only
        self.n_left = int(left.shape[0]) if left is not None else 0
        self.n_own = int(self.xyz.shape[0])
        self.n = sum(int(t.shape[0]) for t in segs)
        if self.n == 0:
            raise ValueError("empty cloud")
        self.requested_cell = float(cell)
        self.x_split = int(x_split)
        h = C.c_void_p()
        ptrs = (C.c_void_p * len(segs))(*[t.data_ptr() for t in segs])
        cnts = (C.c_int64 * len(segs))(*[int(t.shape[0]) for t in segs])
        lat = None if lattice is None else (C.c_double * 2)(float(lattice[0]), float(lattice[1]))
        bnd = None if bounds is None else (C.c_double * 6)(*[float(v) for v in bounds])
        with torch.cuda.device(self.device):
            check(self.lib.more3d_grid_create_segments(ptrs, cnts, len(segs), float(cell), int(x_split), lat, bnd, stream_ptr(),
                                                       C.byref(h)))
        self._h = h
        info = (C.c_int64 * 3)()
        geom = (C.c_double * 3)()
        check(self.lib.more3d_grid_info(self._h, info, geom))
        self.nx, self.ny = int(info[1]), int(info[2])
        self.cell, self.ox, self.oy = float(geom[0]), float(geom[1]), float(geom[2])

    def close(self):
        if getattr(self, "_h", None):
            self.lib.more3d_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------------------------------
    def _new(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def order(self):
        o = self._new((self.n,), torch.int32)
        check(self.lib.more3d_grid_order(self._h, ptr(o), stream_ptr()))
        return o

    def radius_csr(self, q, r, idx_dtype=torch.int32):
        """CSR neighbour lists: (offsets int64 [nq+1], idx [total]).  q=None -> the cloud itself."""
        with torch.cuda.device(self.device):
            qd = None if q is None else as_device_points(q, self.device)
            nq = self.n if qd is None else int(qd.shape[0])
            if nq == 0:     # an empty tensor has a NULL data pointer, which the C-ABI reads as "self query"
                return torch.zeros(1, dtype=torch.int64, device=self.device), self._new((0,), idx_dtype)
            counts = self._new((max(nq, 1),), torch.int32)
            offsets = self._new((nq + 1,), torch.int64)
            check(self.lib.more3d_radius_count(self._h, ptr(qd), nq, float(r), ptr(counts), stream_ptr()))
            total = C.c_int64(0)
            check(self.lib.more3d_offsets_from_counts(ptr(counts), nq, ptr(offsets), C.byref(total), stream_ptr()))
            idx = self._new((max(int(total.value), 1),), idx_dtype)
            nbytes = 4 if idx_dtype == torch.int32 else 8
            check(self.lib.more3d_radius_fill(self._h, ptr(qd), nq, float(r), ptr(offsets), ptr(idx), nbytes,
                                              stream_ptr()))
            return offsets, idx[:int(total.value)]

    def radius_count(self, q, r):
        with torch.cuda.device(self.device):
            qd = None if q is None else as_device_points(q, self.device)
            nq = self.n if qd is None else int(qd.shape[0])
            if nq == 0:
                return self._new((0,), torch.int32)
            counts = self._new((max(nq, 1),), torch.int32)
            check(self.lib.more3d_radius_count(self._h, ptr(qd), nq, float(r), ptr(counts), stream_ptr()))
            return counts[:nq]

    def knn(self, q, k, return_distance=True):
        with torch.cuda.device(self.device):
            qd = None if q is None else as_device_points(q, self.device)
            nq = self.n if qd is None else int(qd.shape[0])
            idx = self._new((nq, k), torch.int64)
            dist = self._new((nq, k), torch.float64) if return_distance else None
            if nq == 0:
                return idx, dist
            check(self.lib.more3d_knn(self._h, ptr(qd), nq, int(k), ptr(idx), ptr(dist), stream_ptr()))
            return idx, dist

    def normals_pca(self, r):
        with torch.cuda.device(self.device):
            normals = self._new((self.n, 3), torch.float64)
            ratio = self._new((self.n,), torch.float64)
            counts = self._new((self.n,), torch.int32)
            check(self.lib.more3d_normals_pca(self._h, float(r), ptr(normals), ptr(ratio), ptr(counts), stream_ptr()))
            return normals, ratio, counts

    def nonparam_curvature(self, r, normals, eps, min_points):
        """normals (N,3) f64 device tensor, flipped in place where the reference flips."""
        with torch.cuda.device(self.device):
            K = self._new((self.n,), torch.float32)
            pcount = self._new((self.n,), torch.int32)
            check(self.lib.more3d_nonparam_curvature(self._h, float(r), ptr(normals), float(eps), int(min_points),
                                                     ptr(K), ptr(pcount), stream_ptr()))
            self._feat_key = None       # normals may have been flipped in place
            return K, pcount

    def normals_curvature(self, r, eps, min_points):
        with torch.cuda.device(self.device):
            normals = self._new((self.n, 3), torch.float64)
            ratio = self._new((self.n,), torch.float64)
            K = self._new((self.n,), torch.float32)
            pcount = self._new((self.n,), torch.int32)
            check(self.lib.more3d_normals_curvature(self._h, float(r), float(eps), int(min_points), ptr(normals),
                                                    ptr(ratio), ptr(K), ptr(pcount), stream_ptr()))
            # the grid now caches the sorted (xyz, K, unit normal) records of exactly these two tensors
            self._feat_key = (normals, normals._version, K, K._version)
            return normals, ratio, K, pcount

    # -- grid-order variants: outputs stay in the grid's sorted order (coalesced writes) -------------------
    def normals_curvature_gridorder(self, r, eps, min_points):
        """Like normals_curvature, but row i of every output belongs to original point order()[i].
        Returns GridColumn objects (lazy un-permutation)."""
        with torch.cuda.device(self.device):
            normals = self._new((self.n, 3), torch.float64)
            ratio = self._new((self.n,), torch.float64)
            K = self._new((self.n,), torch.float32)
            pcount = self._new((self.n,), torch.int32)
            check(self.lib.more3d_normals_curvature_gridorder(self._h, float(r), float(eps), int(min_points),
                                                              ptr(normals), ptr(ratio), ptr(K), ptr(pcount),
                                                              stream_ptr()))
            cols = tuple(GridColumn(self, t) for t in (normals, ratio, K, pcount))
            self._feat_key = None
            self._feat_cols = (cols[0], cols[2])     # the cached feature records hold exactly these two
            return cols

    def dk_dn_gridorder(self, r, rho, sigma_normal, sigma_curvature, chi2_table, min_points, own_lo=0, n_owned=0,
                        status=None, minmax=None):
        """dk / dn from the feature records cached by the last normals_curvature[_gridorder]; GridColumns.
        n_owned > 0: minmax covers the points with original index in [own_lo, own_lo + n_owned) only (a strip's own
        points between its ghosts).  status: device int32 word for the chi2-table overflow (no synchronisation; see
        more3d_dk_dn) -- None makes the call itself check and raise.  minmax: optional (4,) f32 device slice to fill."""
        with torch.cuda.device(self.device):
            dk = self._new((self.n,), torch.float32)
            dn = self._new((self.n,), torch.float32)
            if minmax is None:
                minmax = self._new((4,), torch.float32)
            check(self.lib.more3d_dk_dn_gridorder(self._h, float(r), float(rho), float(sigma_normal),
                                                  float(sigma_curvature), ptr(chi2_table), int(chi2_table.numel()),
                                                  int(min_points), int(own_lo), int(n_owned), ptr(dk), ptr(dn),
                                                  ptr(minmax), ptr(status), stream_ptr()))
            return GridColumn(self, dk), GridColumn(self, dn), minmax

    def order_tensor(self):
        """order() cached on the grid (sorted position -> original index, int32)."""
        o = getattr(self, "_order", None)
        if o is None:
            with torch.cuda.device(self.device):
                o = self._order = self.order()
        return o

    def dk_dn(self, r, normals, K, rho, sigma_normal, sigma_curvature, chi2_table, min_points, status=None):
        """chi2_table: device f64 tensor, chi2.ppf(alpha, df) for df = 0..len-1."""
        with torch.cuda.device(self.device):
            dk = self._new((self.n,), torch.float32)
            dn = self._new((self.n,), torch.float32)
            minmax = self._new((4,), torch.float32)
            key = getattr(self, "_feat_key", None)
            if key is not None and key[0] is normals and key[2] is K and key[1] == normals._version \
                    and key[3] == K._version:
                normals = K = None      # untouched since the fused pass: use the grid's cached feature records
            check(self.lib.more3d_dk_dn(self._h, float(r), ptr(normals), ptr(K), float(rho), float(sigma_normal),
                                        float(sigma_curvature), ptr(chi2_table), int(chi2_table.numel()),
                                        int(min_points), ptr(dk), ptr(dn), ptr(minmax), ptr(status), stream_ptr()))
            return dk, dn, minmax
