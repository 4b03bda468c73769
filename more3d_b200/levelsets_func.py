"""Drop-in for the numerical part of the reference's ``levelsets_func.py`` (point-based Chan-Vese level
set on a kNN graph with WLS derivatives) on the B200 kernels.

Same names, signatures and defaults: ``wendland, normalize, clip, compute_normals,
compute_tangent_planes, build_neighborhoods, Derivative, heaviside, delta, dp, smooth, Solver``
(levelsets_func.py:92-667).  Arrays go in and come out as numpy arrays exactly like the reference;
``Solver`` keeps its state in HBM between calls and mirrors it to the host lazily (only when an attribute
such as ``solver.phi`` is touched).  Both active-contour models run on the device: ``'chan_vese'``
(:386-437) and ``'lmv'`` (:439-513, single-channel cue -- see ``Solver.run``).  The LAS / PLY readers
(:43-84) live in ``more3d_b200/io_formats.py``.

Reference quirks kept on purpose: normals are PCA of offsets about the QUERY point
(:116); ``compute_tangent_planes`` normalises the caller's normals in place (:140); after
``initialize_from_voxels / seeds / point`` ``last_active`` IS ``active`` (:581, :589, :596), which makes
the re-clamping of (de)activated points inert for those initialisations; ``verbose`` is a module global.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .grid import UniformGrid
from .io_formats import read_las, read_ply  # noqa: F401  (levelsets_func.py:43-84)

verbose = False


def _print(string, l=0, **kwargs):
    if verbose:
        print(''.join([' '] * l) + string, **kwargs, **{'flush': True})


def _dev(a, dtype=torch.float64):
    _lib.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype).contiguous()


def _elementwise(op, x, param):
    lib = _lib.load()
    xs = np.asarray(x, dtype=np.float64)
    d = _dev(xs.reshape(-1))
    if d.numel() == 0:
        return np.zeros_like(xs)
    out = torch.empty_like(d)
    check(lib.more3d_ls_elementwise(op, ptr(d), float(param), int(d.numel()), ptr(out), stream_ptr()))
    return out.cpu().numpy().reshape(xs.shape)


def wendland(x, h):
    """wendland weighting function (levelsets_func.py:92-97)"""
    return _elementwise(3, x, h)


def heaviside(x, epsilon=1.0):
    """mollified heaviside function (:319-321)"""
    return _elementwise(0, x, epsilon)


def delta(x, epsilon=1.0):
    """mollified dirac delta (:332-333)"""
    return _elementwise(1, x, epsilon)


def dp(s):
    """dp(s)=d'(s)/s where d is the double-well potential (:336-342)"""
    return _elementwise(2, s, 0.0)


def _percentile_plan(n, q):
    """numpy's `linear` percentile (lib/_function_base_impl.py: virtual index (n - 1) * q / 100, gamma = its
    fractional part): returns (lower rank j, upper rank, interpolation weight gamma)."""
    quant = np.true_divide(q, 100.0)
    vi = (n - 1) * quant
    j = int(np.floor(vi))
    j = min(max(j, 0), n - 1)
    gamma = float(vi - j)
    return j, min(j + 1, n - 1), gamma


def _percentile_bounds_dev(d, q):
    """device tensor (2,) = {0, np.percentile(d, q)} for a flat f64 device tensor, no host round trip"""
    lib = _lib.load()
    n = int(d.numel())
    j, j1, gamma = _percentile_plan(n, q)
    with torch.cuda.device(d.device):
        ab = torch.empty(2, dtype=torch.float64, device=d.device)
        ranks = (C.c_int64 * 2)(j, j1)
        check(lib.more3d_select_kth(ptr(d), n, ranks, 2, ptr(ab), stream_ptr()))
        lohi = torch.empty(2, dtype=torch.float64, device=d.device)
        check(lib.more3d_cue_ops(3, None, 0, ptr(ab), float(gamma), ptr(lohi), stream_ptr()))
    return lohi


def percentile(array, q):
    """np.percentile(array, q) (scalar q, default `linear` method) computed on the device: exact order statistics by
    radix select + numpy's interpolation rule (the clip bound of run_levelsets.py:220)."""
    a = np.asarray(array, dtype=np.float64)
    d = _dev(a.reshape(-1))
    return float(_percentile_bounds_dev(d, q)[1].item())


def _inplace_device(array, fn):
    """run fn(flat f64 device tensor) and write the result back into the caller's numpy array (the reference's
    cue helpers work in place on host arrays; the arithmetic happens on the GPU)"""
    a = np.asarray(array)
    if a.size == 0:
        return
    d = _dev(np.ascontiguousarray(a, dtype=np.float64).reshape(-1))
    fn(d)
    array[...] = d.cpu().numpy().reshape(a.shape).astype(a.dtype, copy=False)


def normalize(array, s=1):
    """in place: shift to a zero minimum, scale to a maximum of s; an all-zero array is left alone (:100-104)"""
    lib = _lib.load()
    _inplace_device(array, lambda d: check(lib.more3d_cue_ops(2, ptr(d), int(d.numel()), None, float(s), None, stream_ptr())))


def clip(array, a, b):
    """in place: array[array < a] = a; array[array > b] = b (:107-109)"""
    lib = _lib.load()

    def run(d):
        lohi = torch.tensor([float(a), float(b)], dtype=torch.float64).to(d.device)
        check(lib.more3d_cue_ops(1, ptr(d), int(d.numel()), ptr(lohi), 0.0, None, stream_ptr()))
    _inplace_device(array, run)


def _mask_of(idx, n):
    """reference index expressions (None / slice(None) / bool mask / int array) -> (uint8 mask or None, selector)"""
    if idx is None or (isinstance(idx, slice) and idx == slice(None)):
        return None, slice(None)
    idx = np.asarray(idx)
    if idx.dtype == bool:
        return idx.astype(np.uint8), idx
    m = np.zeros(n, np.uint8)
    m[idx] = 1
    return m, idx


def smooth(u, idx, nbs, w):
    """kernel smoother with weights w (:350-358): u (N,) or (N,C); returns u[idx]-shaped smoothed values"""
    lib = _lib.load()
    u = np.asarray(u, dtype=np.float64)
    n = u.shape[0]
    cols = 1 if u.ndim == 1 else int(np.prod(u.shape[1:]))
    mask, sel = _mask_of(idx, n)
    du = _dev(u.reshape(n, cols))
    dn = _dev(nbs, torch.int64)
    dw = _dev(w)
    dm = None if mask is None else _dev(mask, torch.uint8)
    out = torch.empty_like(du)
    check(lib.more3d_ls_smooth_raw(ptr(du), cols, ptr(dn), ptr(dw), n, int(dn.shape[1]), ptr(dm), ptr(out), stream_ptr()))
    return out.cpu().numpy().reshape(u.shape)[sel]


def _knn_grid(dpts, k):
    b = _lib.cloud_bounds(dpts)
    ext = b[3:] - b[:3]
    area = max(float(ext[0]) * float(ext[1]), 1e-300)
    cell = float(np.sqrt(max(k, 8) * area / (4.0 * dpts.shape[0])))
    if not np.isfinite(cell) or cell <= 0:
        cell = 1.0
    return UniformGrid(dpts, cell)


def _frames(dpts, knn2k, k, normals):
    lib = _lib.load()
    n = int(dpts.shape[0])
    has = normals is not None
    dn = _dev(normals).clone() if has else torch.empty((n, 3), dtype=torch.float64, device=dpts.device)
    t1 = torch.empty((n, 3), dtype=torch.float64, device=dpts.device)
    t2 = torch.empty((n, 3), dtype=torch.float64, device=dpts.device)
    nb = torch.empty((n, k), dtype=torch.int32, device=dpts.device)
    check(lib.more3d_ls_frames(ptr(dpts), n, int(k), ptr(knn2k), 1 if has else 0, ptr(dn), ptr(t1), ptr(t2), ptr(nb),
                               stream_ptr()))
    return dn, t1, t2, nb


def compute_normals(points, neighbors):
    """compute normals via PCA of neighbors (:112-129); neighbors: (N, m) index array (m <= 32)"""
    neighbors = np.asarray(neighbors)
    if neighbors.ndim != 2:
        raise AssertionError   # the ragged-list branch asserts len(nbs) > 3 (:122-123); only the array form is supported
    n, m = neighbors.shape
    if m % 2:                  # the kernel walks an even number of columns: pad with the point itself (zero offset)
        neighbors = np.column_stack([neighbors, np.arange(n)])
        m += 1
    if m > 32:
        raise ValueError("more than 32 neighbours per point are not supported")
    dn, _, _, _ = _frames(_dev(points), _dev(neighbors, torch.int64), m // 2, None)
    return dn.cpu().numpy()


def compute_tangent_planes(points, neighbors, normals):
    """compute 2 orthogonal vectors spanning the tangent plane at each point (:132-151).
    Like the reference this normalises ``normals`` in place."""
    neighbors = np.asarray(neighbors)
    d = _dev(points)
    m = neighbors.shape[1]
    # only the last column (farthest neighbour) and the normals matter here
    knn = _dev(np.column_stack([neighbors[:, :1], neighbors[:, -1:]]), torch.int64)
    dn, t1, t2, _ = _frames(d, knn, 1, normals)
    if isinstance(normals, np.ndarray):
        normals[...] = dn.cpu().numpy()
    return t1.cpu().numpy(), t2.cpu().numpy()


def build_neighborhoods(points, h, k, normals=None, return_device=False):
    """build point neighborhoods, normals & tangents (:154-172)"""
    _print('building point neighborhoods')
    d = _dev(points)
    n = int(d.shape[0])
    g = _knn_grid(d, 2 * k)
    _print('querying neighbors', 1)
    knn, _ = g.knn(None, 2 * k, return_distance=False)     # tree.query(points, 2k, sort_results=True), :161-163
    g.close()
    _print('computing normals / tangents', 1)
    dn, t1, t2, nb = _frames(d, knn, k, normals)
    _print('neighborhoods done')
    if return_device:
        return nb, dn, (t1, t2)
    return nb.cpu().numpy().astype(np.int64), dn.cpu().numpy(), (t1.cpu().numpy(), t2.cpu().numpy())


class Derivative:
    def __init__(self, points, max_d, neighbors, tangents, weights=None):
        """precalculate inverse for WLS approximation via SVD (:246-275).
        ``weights`` (extension): (N, k) array used instead of the computed Wendland weights -- numpy's array
        power makes the reference's own weights host-CPU dependent in the last bit, and ``smooth`` decides ties
        at the cap with them, so reproducing a given reference run bit for bit needs its weights."""
        lib = _lib.load()
        pts = _dev(points)
        n = int(pts.shape[0])
        nb = neighbors if isinstance(neighbors, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(neighbors))
        nb = nb.to(device="cuda", dtype=torch.int64)
        # Internal point order = the uniform grid's sorted order: a warp's 32 points then share their
        # neighbours' cache lines, which turns the k random 32-B sector gathers per point of every iteration
        # kernel into L1/L2 hits.  `_perm[i]` = original index of internal point i; everything that leaves
        # this class is mapped back, so callers never see the permutation.
        g = _knn_grid(pts, 8)
        self._perm = g.order().long()
        g.close()
        self._inv = torch.empty_like(self._perm)
        self._inv[self._perm] = torch.arange(n, device="cuda", dtype=torch.int64)
        self._pts = pts[self._perm].contiguous()
        self._nb = self._inv[nb[self._perm]].to(torch.int32).contiguous()
        self._k = int(self._nb.shape[1])
        self._t1, self._t2 = _dev(tangents[0])[self._perm].contiguous(), _dev(tangents[1])[self._perm].contiguous()
        given = weights is not None
        self._w = _dev(weights)[self._perm].contiguous() if given else torch.empty((n, self._k), dtype=torch.float64, device="cuda")
        self._D = None
        self._max_d = float(max_d)
        self._host = {}
        h = C.c_void_p()
        sing = C.c_int(0)
        check(lib.more3d_ls_graph_create(ptr(self._pts), n, self._k, ptr(self._nb), ptr(self._t1), ptr(self._t2),
                                         self._max_d, 1 if given else 0, ptr(self._w), None, C.byref(sing), stream_ptr(),
                                         C.byref(h)))
        self._g = h
        self.singular = bool(sing.value)       # the reference opens pdb here (:269-271); we flag it instead
        self._neighbors_in, self._points_in, self._tangents_in = neighbors, points, tangents

    def __del__(self):
        try:
            if getattr(self, "_g", None):
                _lib.load().more3d_ls_graph_destroy(self._g)
                self._g = None
        except Exception:
            pass

    # reference attributes, mirrored to the host on first use
    def _h(self, name, fn):
        if name not in self._host:
            self._host[name] = fn()
        return self._host[name]

    @property
    def weights(self):
        return self._h("weights", lambda: self._w[self._inv].cpu().numpy())

    @property
    def neighbors(self):
        return self._h("neighbors", lambda: self._perm[self._nb.long()][self._inv].cpu().numpy())

    @property
    def points(self):
        return self._h("points", lambda: self._pts[self._inv].cpu().numpy())

    @property
    def t1(self):
        return self._h("t1", lambda: self._t1[self._inv].cpu().numpy())

    @property
    def t2(self):
        return self._h("t2", lambda: self._t2[self._inv].cpu().numpy())

    @property
    def D(self):
        def make():
            lib = _lib.load()
            n = int(self._pts.shape[0])
            D = torch.empty((n, 5, self._k + 3), dtype=torch.float64, device="cuda")
            w = self._w.clone()
            h, sing = C.c_void_p(), C.c_int(0)
            check(lib.more3d_ls_graph_create(ptr(self._pts), n, self._k, ptr(self._nb), ptr(self._t1), ptr(self._t2),
                                             self._max_d, 1, ptr(w), ptr(D), C.byref(sing), stream_ptr(), C.byref(h)))
            lib.more3d_ls_graph_destroy(h)
            return D[self._inv].cpu().numpy()
        return self._h("D", make)

    def _diff(self, u):
        lib = _lib.load()
        du = _dev(u)[self._perm].contiguous()
        out = torch.empty((du.shape[0], 4), dtype=torch.float64, device="cuda")
        check(lib.more3d_ls_diff(self._g, ptr(du), None, ptr(out), stream_ptr()))
        return out[self._inv].cpu().numpy()

    def __call__(self, u, idx=None, first_only=False):
        """first & optionally second derivatives of u at idx w.r.t. local tangent basis (:277-291)"""
        if idx is None:
            idx = slice(None)
        a = self._diff(u)[idx]
        if first_only:
            return (a[:, 0], a[:, 1])
        return (a[:, 0], a[:, 1], a[:, 2], a[:, 3])

    def grad(self, u, idx):
        """gradient of u at idx (:293-298)"""
        if idx is None:
            idx = slice(None)
        dx, dy = self(u, idx, first_only=True)
        return dx[:, None] * self.t1[idx] + dy[:, None] * self.t2[idx]

    def div(self, u, idx):
        """divergence of u at idx (:300-305)"""
        dx_ux = self.grad(u[:, 0], idx)[:, 0]
        dy_uy = self.grad(u[:, 1], idx)[:, 1]
        dz_uz = self.grad(u[:, 2], idx)[:, 2]
        return dx_ux + dy_uy + dz_uz


class _Mirror:
    """A device tensor with a lazily synchronised host numpy mirror.  Reading the host side marks it as
    possibly modified (callers mutate ``solver.phi[:] = ...`` in place), so it is pushed back before the
    next device use."""

    def __init__(self, host, perm=None, inv=None):
        self.host = host
        self.dev = None              # internal (grid-sorted) point order when perm is given
        self.perm, self.inv = perm, inv
        self.host_touched = True     # host is authoritative
        self.host_stale = False

    def get_host(self):
        if self.host_stale:
            d = self.dev if self.inv is None else self.dev[self.inv]
            self.host[...] = d.cpu().numpy().astype(self.host.dtype, copy=False).reshape(self.host.shape)
            self.host_stale = False
        self.host_touched = True
        return self.host

    def get_dev(self, dtype):
        if self.dev is None or self.host_touched:
            d = torch.from_numpy(np.ascontiguousarray(self.host)).to(device="cuda", dtype=dtype)
            self.dev = (d if self.perm is None else d[self.perm]).contiguous()
            self.host_touched = False
        return self.dev

    def dev_written(self):
        self.host_stale = True


class Solver:
    def __init__(self, h, zeta, points, neighbors, tangents, active_contour_model='chan_vese', weights=None):
        """level set solver (:362-384); ``weights``: see :class:`Derivative`"""
        _print('initializing solver')
        zeta = np.asarray(zeta.cpu().numpy() if isinstance(zeta, torch.Tensor) else zeta, dtype=np.float64)
        self.h = h
        self.max_val = 2 * h
        _print('initializing WLS derivatives', 1)
        self.diff = Derivative(points, 3 * h, neighbors, tangents, weights=weights)
        self._phi = self._mirror(np.zeros(zeta.shape[:-1]))
        self._zeta = self._mirror(zeta.copy())
        self._active = self._mirror(np.zeros(self._phi.host.shape, dtype='bool'))
        self._last_active = self._mirror(self._active.host.copy())
        self._alias = False
        self.seeds = None
        self.active_contour_model = active_contour_model

    def _mirror(self, host):
        return _Mirror(host, self.diff._perm, self.diff._inv)

    # -- reference attributes ---------------------------------------------------------------------
    @property
    def phi(self):
        return self._phi.get_host()

    @phi.setter
    def phi(self, v):
        self._phi.get_host()[...] = v

    @property
    def zeta(self):
        return self._zeta.get_host()

    @zeta.setter
    def zeta(self, v):
        self._zeta = self._mirror(np.asarray(v, dtype=np.float64))

    @property
    def active(self):
        return self._active.get_host()

    @active.setter
    def active(self, v):
        self._active = self._mirror(np.asarray(v, dtype=bool).copy())

    @property
    def last_active(self):
        return self._active.get_host() if self._alias else self._last_active.get_host()

    @last_active.setter
    def last_active(self, v):
        if v is self._active.host:
            self._alias = True
        else:
            self._alias = False
            self._last_active = self._mirror(np.asarray(v, dtype=bool).copy())

    @property
    def points(self):
        return self.diff.points

    @property
    def neighbors(self):
        return self.diff.neighbors

    @property
    def neighborhood(self):
        return np.column_stack((np.arange(self.points.shape[0]), self.neighbors))

    # -- helpers ------------------------------------------------------------------------------------
    def zero_crossings(self, u, idx=None, neighbors=None):
        """get all points (indices) which have a neighbor with differing sign (:524-540)"""
        if neighbors is not None and neighbors is not self.neighbors:
            raise NotImplementedError("zero_crossings on a neighbour table other than the solver's own")
        lib = _lib.load()
        n = int(self.diff._pts.shape[0])
        mask, sel = _mask_of(idx, n)
        p, inv = self.diff._perm, self.diff._inv
        du = _dev(u)[p].contiguous()
        dm = None if mask is None else _dev(mask, torch.uint8)[p].contiguous()
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        check(lib.more3d_ls_zero_crossings(self.diff._g, ptr(du), ptr(dm), ptr(out), stream_ptr()))
        return out[inv].cpu().numpy().astype(bool)[sel]

    def initialize_new(self, new, active):
        phi = self.phi
        phi[new] = np.sign(phi[new]) * self.max_val
        phi[self.neighbors[(new & ~active)]] = np.sign(phi[(new & ~active), None]) * self.max_val

    def _set_active_from_phi(self):
        lib = _lib.load()
        n = int(self.diff._pts.shape[0])
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        check(lib.more3d_ls_zero_crossings(self.diff._g, ptr(self._phi.get_dev(torch.float64)), None, ptr(out), stream_ptr()))
        m = self._mirror(np.zeros(n, dtype=bool))
        m.dev = out                      # already in internal order
        m.host_touched = False
        m.host_stale = True
        self._active = m

    def initialize_from_mask(self, mask):
        _print('initializing from mask')
        phi = self.phi
        phi[mask] = -self.max_val
        phi[~mask] = self.max_val
        self._set_active_from_phi()
        self._alias = False
        self._last_active = self._mirror(self._active.get_host().copy())   # last_active[:] = active, :553

    def preprocess_cue(self, num_smooth=0, cue_clip_pc=99.9, scale=1):
        """The driver's cue pre-processing (run_levelsets.py:211-221) with the cue resident in HBM throughout:
        `num_smooth` kernel-smoothing passes, zeta = |zeta|, clip(zeta, 0, np.percentile(zeta, cue_clip_pc)),
        normalize(zeta, scale).  Same values as the host sequence ``smooth / np.abs / clip / normalize`` on
        ``solver.zeta``; nothing is copied to the host."""
        lib = _lib.load()
        z = self._zeta.get_dev(torch.float64)
        n = int(z.shape[0])
        cols = 1 if z.ndim == 1 else int(z.shape[1])
        with torch.cuda.device(z.device):
            for _ in range(int(num_smooth)):
                # smooth(zeta, slice(None), neighbors, weights) (:350-358) column by column on the solver's own graph
                cols_in = [z.reshape(n)] if cols == 1 else [z[:, c].contiguous() for c in range(cols)]
                cols_out = []
                for u in cols_in:
                    o = torch.empty_like(u)
                    check(lib.more3d_ls_smooth(self.diff._g, ptr(u), None, ptr(o), stream_ptr()))
                    cols_out.append(o)
                z = cols_out[0].reshape(z.shape) if cols == 1 else torch.stack(cols_out, dim=1).contiguous()
            flat = z.reshape(-1)
            check(lib.more3d_cue_ops(0, ptr(flat), int(flat.numel()), None, 0.0, None, stream_ptr()))
            lohi = _percentile_bounds_dev(flat, cue_clip_pc)
            check(lib.more3d_cue_ops(1, ptr(flat), int(flat.numel()), ptr(lohi), 0.0, None, stream_ptr()))
            check(lib.more3d_cue_ops(2, ptr(flat), int(flat.numel()), None, float(scale), None, stream_ptr()))
        self._zeta.dev = z
        self._zeta.dev_written()

    def initialize_from_zeta(self, percentile):
        _print('initializing from zeta')
        zeta = self.zeta
        pc = np.percentile(zeta, percentile)
        phi = self.phi
        phi[zeta <= pc] = -self.max_val          # IndexError for (N, C) zeta exactly as in the reference (:559)
        phi[zeta > pc] = self.max_val
        self._set_active_from_phi()
        self._alias = False
        self._last_active = self._mirror(self._active.get_host().copy())

    def initialize_from_seeds(self, num_seeds, seed_radius, reuse_seeds=True):
        pts = self.points
        if self.seeds is None or not reuse_seeds:
            _print('generating seeds')
            self.seeds = np.zeros((pts.shape[0]), dtype=bool)
            for _ in range(num_seeds):
                i = np.random.choice(pts.shape[0], 1)[0]
                self.seeds[np.linalg.norm((pts - pts[i]), axis=1) < seed_radius] = True
        else:
            _print('reusing seeds')
        phi = self.phi
        phi[self.seeds] = self.max_val
        phi[~self.seeds] = -self.max_val
        self._set_active_from_phi()
        self._alias = True                        # self.last_active = self.active, :581

    def initialize_from_point(self, point, radius):
        pts = self.points
        i = np.linalg.norm((point - pts), axis=(-1)) <= radius
        phi = self.phi
        phi[~i] = self.max_val
        phi[i] = -self.max_val
        self._set_active_from_phi()
        self._alias = True                        # :589

    def initialize_from_voxels(self, h):
        lib = _lib.load()
        n = int(self.diff._pts.shape[0])
        dphi = torch.empty(n, dtype=torch.float64, device="cuda")
        check(lib.more3d_ls_init_voxels(ptr(self.diff._pts), n, float(h), float(self.max_val), ptr(dphi), stream_ptr()))
        self._phi.dev = dphi
        self._phi.host_touched = False
        self._phi.dev_written()
        self._set_active_from_phi()
        self._alias = True                        # :596

    # -- the iteration ----------------------------------------------------------------------------------
    def run(self, iterations, stepsize, nu, mu, lambda1, lambda2, epsilon, cap=True, tolerance=1e-4):
        """explicit Euler iterations of the Chan-Vese or local-mean-and-variance flow (:598-642).  Returns True
        when the RMS change drops below ``tolerance``."""
        _print(f"running {iterations} steps")
        assert self.active_contour_model in ['chan_vese', 'lmv']
        lib = _lib.load()
        phi = self._phi.get_dev(torch.float64)
        zeta = self._zeta.get_dev(torch.float64)
        channels = int(zeta.shape[-1]) if zeta.ndim > 1 else 1
        active = self._active.get_dev(torch.uint8)
        last = active if self._alias else self._last_active.get_dev(torch.uint8)
        rmsc = (C.c_double * max(int(iterations), 1))()
        done, conv = C.c_int(0), C.c_int(0)
        if self.active_contour_model == 'lmv':
            # _lmv_step (:439-513) is written for a 1-D cue: with the (N, C) cue the constructor asks for, its
            # zeta[nb] * (1 - Hphi[nb]) raises ValueError (shapes (a, k+1, C) x (a, k+1)), and it runs verbatim
            # once the caller sets solver.zeta = solver.zeta[:, 0].  Here a single-channel cue runs either way
            # ((N, 1) is squeezed); more channels raise like the reference.
            if channels != 1:
                raise ValueError("operands could not be broadcast together: the 'lmv' model takes a "
                                 "single-channel cue, got %d channels" % channels)
            check(lib.more3d_ls_run_lmv(self.diff._g, ptr(phi), ptr(zeta), ptr(active), ptr(last),
                                        1 if self._alias else 0, int(iterations), float(stepsize), float(nu), float(mu),
                                        float(epsilon * self.h), float(self.max_val), 1 if cap else 0, float(tolerance),
                                        rmsc, C.byref(done), C.byref(conv), stream_ptr()))
        else:
            check(lib.more3d_ls_run(self.diff._g, ptr(phi), ptr(zeta), channels, ptr(active), ptr(last), 1 if self._alias else 0,
                                    int(iterations), float(stepsize), float(nu), float(mu), float(lambda1), float(lambda2),
                                    float(epsilon * self.h), float(self.max_val), 1 if cap else 0, float(tolerance), rmsc,
                                    C.byref(done), C.byref(conv), stream_ptr()))
        self._phi.dev_written()
        self._active.dev_written()
        if not self._alias:
            self._last_active.dev_written()
        self.rmsc_history = [rmsc[i] for i in range(done.value)]
        if done.value:
            _print(f"\r{done.value}/{iterations} \t RMSE(dphi): {rmsc[done.value - 1]:.3e}", 1, end='')
        if conv.value:
            _print("tolerance reached!")
            return True
        return False

    def extract(self, extraction_threshold=4):
        """the region-growing part of Solver.save(extract=True) (:651-661), without the file output; on the device"""
        lib = _lib.load()
        phi = self._phi.get_dev(torch.float64)
        zeta = self._zeta.get_dev(torch.float64)
        channels = int(zeta.shape[-1]) if zeta.ndim > 1 else 1
        n = int(self.diff._pts.shape[0])
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        check(lib.more3d_ls_extract(self.diff._g, ptr(phi), ptr(zeta), channels, int(extraction_threshold), ptr(out),
                                    stream_ptr()))
        return out[self.diff._inv].cpu().numpy().astype(bool)

    def save(self, basepath, origin=np.r_[0, 0, 0], extract=False, extraction_threshold=4):
        """text output of the active points (and the extracted region), :644-667 -- host-side I/O"""
        act = self.active
        out = np.column_stack((self.points[act] + origin, self.phi[act], self.zeta[act]))
        np.savetxt(basepath + '.txt', out, fmt='%1.6f')
        if extract:
            salient = self.extract(extraction_threshold)
            out = np.column_stack((self.points[salient] + origin[None, :], self.phi[salient], self.zeta[salient]))
            np.savetxt(basepath + '_extract.txt', out, fmt='%1.6f')
