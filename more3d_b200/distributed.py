"""Multi-GPU saliency: the cloud is cut into x-strips, one per rank (one process per GPU), each
holding ghost halos of width 2*r_max from both neighbours.

Data-path collectives (NCCL over NVLink / NVSwitch through torch.distributed):
  * halo exchange of coordinates -- once per cloud, point-to-point with the two strip neighbours;
  * all-reduce(min/max) of the four dk/dn extrema -- once per radius (the Histo pass of
    DM_saliency.py:562-568 must see the whole cloud).
With a 2*r halo every owned point AND every ghost within r of the strip has a complete
neighbourhood, so ghost normals / curvatures are recomputed locally instead of being exchanged.
The communication logic is backend-agnostic (gloo on CPU tensors in the tests).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr, stream_ptr
from .grid import GridColumn, UniformGrid, CELL_MARGIN, X_SPLIT, cell_for_radius


def strip_bounds(x_min, x_max, world):
    """Equal-width strip edges; the last edge is +inf-safe (points exactly on x_max belong to the last strip)."""
    edges = np.linspace(x_min, x_max, world + 1)
    return edges


def band_select(xyz, axis, lo, hi):
    """Stable indices of the device points with lo <= coord < hi (hand-written flag/scan/scatter kernels)."""
    lib = _lib.load()
    n = int(xyz.shape[0])
    idx = torch.empty(n, dtype=torch.int64, device=xyz.device)
    cnt = C.c_int64(0)
    with torch.cuda.device(xyz.device):
        check(lib.more3d_band_select(ptr(xyz), n, int(axis), float(lo), float(hi), ptr(idx), C.byref(cnt), stream_ptr()))
    return idx[:int(cnt.value)]


def gather_rows(src, idx):
    lib = _lib.load()
    m = int(idx.shape[0])
    ncols = int(src.shape[1]) if src.ndim == 2 else 1
    out = torch.empty((m, ncols) if src.ndim == 2 else (m,), dtype=src.dtype, device=src.device)
    with torch.cuda.device(src.device):
        check(lib.more3d_gather_rows(ptr(src), ptr(idx), m, ncols, src.element_size(), ptr(out), stream_ptr()))
    return out


def exchange_halos(send_left, send_right, rank, world, group=None):
    """Send boundary strips to the two x-neighbours and receive theirs.  Tensors are (m, 3) float64 on
    the backend's device.  Returns (from_left, from_right); missing neighbours give empty tensors."""
    dev, dt = send_left.device, send_left.dtype
    left, right = rank - 1, rank + 1
    # 1) sizes
    n_to = torch.tensor([send_left.shape[0], send_right.shape[0]], dtype=torch.int64, device=dev)
    n_from = torch.zeros(2, dtype=torch.int64, device=dev)
    ops = []
    if left >= 0:
        ops += [dist.P2POp(dist.isend, n_to[0:1], left, group), dist.P2POp(dist.irecv, n_from[0:1], left, group)]
    if right < world:
        ops += [dist.P2POp(dist.isend, n_to[1:2], right, group), dist.P2POp(dist.irecv, n_from[1:2], right, group)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    n_from = n_from.cpu()
    from_left = torch.empty((int(n_from[0]), 3), dtype=dt, device=dev)
    from_right = torch.empty((int(n_from[1]), 3), dtype=dt, device=dev)
    # 2) payload
    ops = []
    if left >= 0:
        if send_left.shape[0]:
            ops.append(dist.P2POp(dist.isend, send_left.contiguous(), left, group))
        if from_left.shape[0]:
            ops.append(dist.P2POp(dist.irecv, from_left, left, group))
    if right < world:
        if send_right.shape[0]:
            ops.append(dist.P2POp(dist.isend, send_right.contiguous(), right, group))
        if from_right.shape[0]:
            ops.append(dist.P2POp(dist.irecv, from_right, right, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return from_left, from_right


def allreduce_minmax(minmax, group=None):
    """Global {min_dk, max_dk, min_dn, max_dn} (in place): MIN over slots 0, 2 and MAX over 1, 3."""
    v = minmax.clone()
    v[1] = -v[1]
    v[3] = -v[3]
    dist.all_reduce(v, op=dist.ReduceOp.MIN, group=group)   # one collective: max(x) = -min(-x)
    v[1] = -v[1]
    v[3] = -v[3]
    minmax.copy_(v)
    return minmax


def otsu_from_counts(counts, edges):
    """skimage.filters.threshold_otsu on a (globally reduced) 256-bin histogram -- host side, 256 entries."""
    counts = np.asarray(counts, dtype=np.float64)
    centers = (edges[:-1] + edges[1:]) / 2.0
    with np.errstate(all="ignore"):
        w1 = np.cumsum(counts)
        w2 = np.cumsum(counts[::-1])[::-1]
        m1 = np.cumsum(counts * centers) / w1
        m2 = (np.cumsum((counts * centers)[::-1]) / w2[::-1])[::-1]
        var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return float(centers[int(np.argmax(var12))])


def global_otsu(values, local_histogram, group=None):
    """Otsu threshold of a column that is sharded over the ranks (Downsample.py:54 on the whole cloud):
    all-reduce(min, max) -> common bin edges -> local 256-bin histograms -> all-reduce(sum) -> host scan.
    values: this rank's float32 / float64 tensor; local_histogram(values, edges) -> int64 tensor of 256 counts
    (the CUDA kernel on GPUs, numpy in the gloo tests)."""
    lo_hi = torch.stack([values.min().double(), -values.max().double()])
    dist.all_reduce(lo_hi, op=dist.ReduceOp.MIN, group=group)
    lo, hi = float(lo_hi[0]), -float(lo_hi[1])
    if lo == hi:
        return lo
    edges = np.linspace(lo, hi, 257)
    hist = local_histogram(values, edges)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return otsu_from_counts(hist.cpu().numpy(), edges)


class DeviceOtsu:
    """global_otsu for one or several sharded columns without host round trips between its steps: extrema (one
    min/max kernel per column), ONE all-reduce for all columns, bin edges with numpy.linspace's arithmetic evaluated
    on the device (arange * step + start, last edge = stop: separate multiply and add, so no FMA), local histograms,
    ONE all-reduce.  Nothing synchronises until result(): a strip queues the Otsu of all its radii behind the
    saliency kernels and reads the thresholds with one wait."""

    def __init__(self, values, group=None):
        lib = _lib.load()
        self.single = isinstance(values, torch.Tensor)
        cols = [values.contiguous()] if self.single else [v.contiguous() for v in values]
        dev = cols[0].device
        m = len(cols)
        with torch.cuda.device(dev):
            mm = torch.empty((m, 4), dtype=torch.float32, device=dev)
            for i, v in enumerate(cols):
                check(lib.more3d_minmax4(ptr(v), ptr(v), int(v.numel()), ptr(mm[i]), stream_ptr()))
            lo_hi = torch.stack([mm[:, 0].double(), -mm[:, 1].double()], dim=1)          # (m, 2)
            dist.all_reduce(lo_hi, op=dist.ReduceOp.MIN, group=group)
            lo, hi = lo_hi[:, 0:1], -lo_hi[:, 1:2]
            step = (hi - lo) / 256.0
            edges = torch.arange(257, dtype=torch.float64, device=dev)[None, :] * step
            edges = edges + lo
            edges[:, 256] = hi[:, 0]
            edges = edges.contiguous()
            self.hist = torch.empty((m, 256), dtype=torch.int64, device=dev)
            for i, v in enumerate(cols):
                check(lib.more3d_histogram256(ptr(v), v.element_size(), int(v.numel()), ptr(edges[i]), ptr(self.hist[i]),
                                              stream_ptr()))
            dist.all_reduce(self.hist, op=dist.ReduceOp.SUM, group=group)
            self.packed = torch.cat([lo, hi, edges], dim=1)                              # (m, 259)

    def results(self):
        p = self.packed.cpu().numpy()
        h = self.hist.cpu().numpy()
        out = []
        for i in range(p.shape[0]):
            lo, hi, edges = float(p[i, 0]), float(p[i, 1]), p[i, 2:]
            out.append(lo if lo == hi else otsu_from_counts(h[i], edges))
        return out

    def result(self):
        r = self.results()
        return r[0] if self.single else r


def cuda_histogram(values, edges):
    lib = _lib.load()
    v = values.contiguous()
    with torch.cuda.device(v.device):
        de = edges if isinstance(edges, torch.Tensor) else torch.from_numpy(edges).to(v.device)
        hist = torch.empty(256, dtype=torch.int64, device=v.device)
        check(lib.more3d_histogram256(ptr(v), v.element_size(), int(v.numel()), ptr(de), ptr(hist), stream_ptr()))
    return hist


class _OwnedRows:
    """The owned rows of a grid-order column, brought back to point order only when read."""

    def __init__(self, col, lo, n):
        self.col, self.lo, self.n = col, lo, n

    def tensor(self):
        return self.col.tensor()[self.lo:self.lo + self.n]


def strip_pack(own, left_below, right_from, cap):
    """Both halo bands of a strip and its bounding box in ONE pass, nothing synchronised (more3d_strip_pack):
    returns (left (cap, 3), right (cap, 3), info (8,) f64 = bbox6 + the two row counts), all on the device."""
    lib = _lib.load()
    dev = own.device
    with torch.cuda.device(dev):
        left = torch.empty((cap, 3), dtype=torch.float64, device=dev)
        right = torch.empty((cap, 3), dtype=torch.float64, device=dev)
        info = torch.empty(8, dtype=torch.float64, device=dev)
        check(lib.more3d_strip_pack(ptr(own), int(own.shape[0]), float(left_below), float(right_from), ptr(left),
                                    ptr(right), int(cap), ptr(info), stream_ptr()))
    return left, right, info


class StripStep:
    """The multi-radius saliency of ONE strip, in phases, so that the caller places the collectives between them:

        step = StripStep(own, left, right, radii, params, lattice, bounds)
        step.features()            # grid + per radius: normals / curvature -> dk / dn;  step.minmax (R, 4) local extrema
        ... all-reduce step.minmax (MIN / MAX) ...
        step.blend(weight)         # normalise + blend with the GLOBAL extrema;          step.sal_minmax (R, 2) local
        ... all-reduce step.sal_minmax ...
        step.histograms()          # 256-bin histograms over the global range;           step.hist (R, 256) local
        ... all-reduce step.hist (SUM) ...

    One grid serves all radii; its cloud is [left ghosts, own points, right ghosts] indexed in place; with the
    tile's `lattice` every per-point result is bit-identical to the single-grid result of the whole tile.  Nothing
    in here synchronises the host except the grid build's column-population read-back."""

    def __init__(self, own, left, right, radii, params, lattice=None, bounds=None, chi2_table_len=1024):
        self.radii = tuple(sorted(radii))
        self.params = dict(params)
        self.dev = own.device
        self.n_own = int(own.shape[0])
        self.n_left = 0 if left is None else int(left.shape[0])
        self.own, self.left, self.right = own, left, right
        self.lattice, self.bounds = lattice, bounds
        self.table_len = int(chi2_table_len)
        R = len(self.radii)
        self.minmax = torch.empty((R, 4), dtype=torch.float32, device=self.dev)
        self.sal_minmax = torch.empty((R, 2), dtype=torch.float32, device=self.dev)   # rows {min, max} of the owned saliency
        from .DM_saliency import _status_words
        self.status = _status_words(self.dev, R)
        self.hist = None
        self.grid = None

    def build_grid(self):
        from .DM_saliency import Curvature_Kernel, _chi2_table
        p = self.params
        ck = Curvature_Kernel()
        ck.setArgs(alpha=p["alpha"], roughness=p["roughness"], min_points=p["min_points"])
        self._eps = ck.epsilon()
        self.grid = UniformGrid(self.own, cell_for_radius(self.radii[0]), left=self.left, right=self.right,
                                lattice=self.lattice, bounds=self.bounds, x_split=X_SPLIT)
        self._table = _chi2_table(p["alpha"], self.table_len, self.dev)
        self.dk, self.dn, self.pcount, self.sal_grid = {}, {}, {}, {}

    def features_radius(self, i):
        """normals / curvature -> dk / dn of radius i (grid order); fills self.minmax[i] with the owned rows' extrema."""
        p, g, r = self.params, self.grid, self.radii[i]
        if cell_for_radius(r) / g.requested_cell > 2.05:
            raise ValueError("radii of one strip step must lie within a factor 2 (one shared grid)")
        rho = r / np.sqrt(2.0)
        _, _, _, pcount = g.normals_curvature_gridorder(r, self._eps, p["min_points"])
        dk, dn, _ = g.dk_dn_gridorder(r, rho, p["sigma_normal"], p["sigma_curvature"], self._table, p["min_points"],
                                      own_lo=self.n_left, n_owned=self.n_own, status=self.status[i:i + 1],
                                      minmax=self.minmax[i])
        self.dk[r], self.dn[r] = dk, dn
        self.pcount[r] = _OwnedRows(pcount, self.n_left, self.n_own)

    def features(self):
        self.build_grid()
        for i in range(len(self.radii)):
            self.features_radius(i)
        return self.minmax

    def blend_radius(self, i, curvature_weight):
        """normalise + blend radius i with the (global) extrema in self.minmax[i]; fills self.sal_minmax[i] with the
        extrema of the OWNED rows."""
        lib = _lib.load()
        r, n = self.radii[i], self.grid.n
        with torch.cuda.device(self.dev):
            sal = torch.empty(n, dtype=torch.float32, device=self.dev)      # blend is elementwise: grid order
            # ghosts are other strips' points (and near the halo's outer edge their neighbourhoods are incomplete)
            check(lib.more3d_saliency_blend_owned(self.grid._h, ptr(self.dk[r].data), ptr(self.dn[r].data),
                                                  ptr(self.minmax[i]), float(curvature_weight), self.n_left,
                                                  self.n_own, ptr(sal), ptr(self.sal_minmax[i]), stream_ptr()))
            self.sal_grid[r] = GridColumn(self.dk[r], sal)

    def blend(self, curvature_weight):
        """normalise + blend every radius with the (global) extrema in self.minmax; fills self.sal_minmax with the
        extrema of the OWNED rows (rows {min, max}: ready for the MIN all-reduce)."""
        for i in range(len(self.radii)):
            self.blend_radius(i, curvature_weight)
        return self.sal_minmax

    def saliency(self, r):
        """(n_own,) f32 saliency of the owned points, in the caller's point order."""
        return self.sal_grid[r].tensor()[self.n_left:self.n_left + self.n_own]

    def owned_minmax(self):
        """Local extrema of the owned points' saliency, rows {min, max} f32 (filled by blend) -- input of the second
        all-reduce."""
        return self.sal_minmax

    def histograms(self):
        """256-bin histograms of the owned saliency over the (global) ranges in sal_minmax, (R, 256) int64."""
        lib = _lib.load()
        R = len(self.radii)
        with torch.cuda.device(self.dev):
            self.hist = torch.empty((R, 256), dtype=torch.int64, device=self.dev)
            for i, r in enumerate(self.radii):
                v = self.saliency(r)
                check(lib.more3d_histogram256_range(ptr(v), 4, self.n_own, ptr(self.sal_minmax[i]), 4, ptr(self.hist[i]),
                                                    stream_ptr()))
        return self.hist

    def result(self):
        """What outlives the step: the overflow words and the Otsu inputs (a few KB)."""
        return _StepResult(self.radii, self.status, self.table_len, self.sal_minmax, self.hist)

    def overflow(self):
        return self.result().overflow()

    def close(self):
        if self.grid is not None:
            self.grid.close()
            self.grid = None


class _StepResult:
    def __init__(self, radii, status, table_len, sal_minmax, hist):
        self.radii, self.status, self.table_len, self.sal_minmax, self.hist = radii, status, table_len, sal_minmax, hist

    def overflow(self):
        """0, or the chi2-table length that would have sufficed (reads the status words: synchronises)."""
        need = int(_lib.read_back(self.status).max())
        if need:
            from .DM_saliency import _status_ring_poisoned
            _status_ring_poisoned(self.status.device)
        return need if need > self.table_len else 0

    def otsu(self):
        """Per-radius Otsu thresholds from the (globally reduced) histograms: the step's ONE host read."""
        lohi = self.sal_minmax.cpu().numpy().astype(np.float64)
        h = self.hist.cpu().numpy()
        out = {}
        for i, r in enumerate(self.radii):
            lo, hi = float(lohi[i, 0]), float(lohi[i, 1])
            out[r] = lo if lo == hi else otsu_from_counts(h[i], np.linspace(lo, hi, 257))
        return out


def allreduce_minmax_rows(mm, group=None):
    """Rows of {min, max[, min, max]} f32 made global in place with ONE collective: MIN over (min, -max); the sign
    flips are one tiny library kernel each (more3d_negate_odd)."""
    lib = _lib.load()
    with torch.cuda.device(mm.device):
        check(lib.more3d_negate_odd(ptr(mm), int(mm.numel()), stream_ptr()))
        dist.all_reduce(mm, op=dist.ReduceOp.MIN, group=group)
        check(lib.more3d_negate_odd(ptr(mm), int(mm.numel()), stream_ptr()))
    return mm


def strip_saliency(own, ghosts, radii, params, curvature_weight, reduce_minmax=None, out_host=None,
                   copy_stream=None, left=None, lattice=None, bounds=None):
    """Saliency of the OWNED points of one strip.  own: (n,3) device f64; ghosts: (g,3) device f64 = the points
    received from the RIGHT neighbour (or all ghosts, when `left` is None and bit-identity with the whole-tile result
    is not required); left: those from the left neighbour.  reduce_minmax(minmax_row) makes one radius' dk/dn
    extrema global (None for a single strip).  Returns ({r: saliency (n,) f32}, {r: owned _pcount rows})."""
    step = StripStep(own, left, ghosts if ghosts is not None and ghosts.shape[0] else None, radii, params,
                     lattice=lattice, bounds=bounds)
    step.features()
    if reduce_minmax is not None:
        for i, r in enumerate(step.radii):
            reduce_minmax(step.minmax[i])
    step.blend(curvature_weight)
    out = {}
    for r in step.radii:
        sal = step.saliency(r)
        if out_host is not None:
            _copy_out(sal, out_host[r], copy_stream, own.device)
        out[r] = sal
    need = step.overflow()
    pc = step.pcount
    step.close()
    if need:
        raise _lib.More3DError("dk/dn: chi2 table too short (%d needed)" % need, _lib.E_RANGE)
    return out, pc


def _copy_out(sal, host, copy_stream, dev):
    if copy_stream is None:
        host.copy_(sal, non_blocking=True)
        return
    ev = torch.cuda.Event()                     # D2H on a side stream: overlaps the kernels queued next
    ev.record(torch.cuda.current_stream(dev))
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ev)
        host.copy_(sal, non_blocking=True)
    from .pipeline import hold_until_copied
    hold_until_copied(sal, copy_stream)


class StripSaliency:
    """bench.py's multi-GPU job: rank k owns an n-point x-strip of a tile.

    One step = pack both halo bands + bounding box (one kernel) -> ONE all-gather of 8 doubles per rank (global
    bounding box = the shared column lattice, and the neighbours' halo sizes; the step's only host read before the
    grid build) -> point-to-point halo exchange -> StripStep: grid over [left ghosts, own, right ghosts] on the
    tile's lattice, 3 radii of normals / curvature / dk / dn -> ONE all-reduce of all radii's dk/dn extrema (the
    Histo pass of DM_saliency.py:562-568 over the whole tile) -> blend.  Two small collectives per step.  The global
    Otsu threshold of the downsampling stage (Downsample.py:54: saliency extrema, 256-bin histograms, two more
    all-reduces) is computed on demand by `otsu`, as on a single GPU where it is not part of the saliency step."""

    def __init__(self, n, rank, world, device, radii, params, curvature_weight, seed=2, cloud=None, x_range=None,
                 host_buffers=True):
        """Default: rank k generates its own n-point strip on the host (weak scaling).  With `cloud` (an (n, 3) f64
        device tensor holding this rank's strip of a larger tile) and `x_range` = (x_lo, x_hi) the job runs on that
        strip instead (bench.py's strong-scaling leg); `host_buffers=False` skips the pinned e2e buffers."""
        from . import synthetic
        self.rank, self.world, self.dev = rank, world, device
        self.radii, self.params, self.cw = tuple(sorted(radii)), dict(params), curvature_weight
        self.halo = 2.0 * max(radii) * CELL_MARGIN
        if cloud is None:
            W = float(np.sqrt(n / synthetic.DENSITY))
            self.W = W
            pts = synthetic.terrain(n, seed=seed + rank, extent=(W, W), origin=(rank * W, 0.0),
                                    centre=(0.5 * world * W, 0.5 * W))
            self.x_lo = rank * W - 0.5 * world * W
            self.x_hi = self.x_lo + W
            self.host = torch.from_numpy(pts).pin_memory()
            self.resident = self.host.to(device)
        else:
            self.resident = cloud
            n = int(cloud.shape[0])
            self.x_lo, self.x_hi = float(x_range[0]), float(x_range[1])
            self.host = None
        self.n = n
        if host_buffers and self.host is not None:
            self.out_host = {r: torch.empty(n, dtype=torch.float32).pin_memory() for r in radii}
            self.h2d_bytes = self.host.numel() * 8
            self.d2h_bytes = n * 4 * len(radii)
        else:
            self.out_host, self.h2d_bytes, self.d2h_bytes = None, 0, 0
        # halo capacity: expected band population from the strip's mean density, with slack; a band that does not
        # fit raises (the count is checked every step)
        self.cap = int(1.5 * n * self.halo / max(self.x_hi - self.x_lo, 1e-9)) + 4096
        self._last = None
        out, pc = self._run(self.resident, None)
        self.sum_m = {r: float(pc[r].tensor().double().sum()) for r in self.radii}
        # partition-independent checksums of the owned points: sum of the neighbour counts and of the saliency
        # columns' float32 BIT PATTERNS (mod 2^64) -- equal across 1/2/4/8-GPU partitions iff every value is
        self.checksum = {r: (int(pc[r].tensor().long().sum()), int(out[r].view(torch.int32).long().sum())) for r in self.radii}

    def _run(self, own, out_host, copy_stream=None):
        rank, world = self.rank, self.world
        left_below = self.x_lo + self.halo if rank > 0 else -np.inf
        right_from = self.x_hi - self.halo if rank < world - 1 else np.inf
        send_l, send_r, info = strip_pack(own, left_below, right_from, self.cap)
        allinfo = torch.empty((world, 8), dtype=torch.float64, device=self.dev)
        dist.all_gather_into_tensor(allinfo, info)
        # the step's host read: global box + every halo size.  Not `.cpu()`: a small copy would queue in the download
        # engine behind the previous step's saliency columns (120 MB) and stall the whole step for milliseconds
        ai = _lib.read_back(allinfo)
        cnt = ai[:, 6:8].astype(np.int64)
        if cnt.max() > self.cap:
            raise _lib.More3DError("halo band holds %d points, capacity %d" % (int(cnt.max()), self.cap), _lib.E_RANGE)
        gmin, gmax = ai[:, 0:3].min(axis=0), ai[:, 3:6].max(axis=0)
        n_from_l = int(cnt[rank - 1, 1]) if rank > 0 else 0
        n_from_r = int(cnt[rank + 1, 0]) if rank < world - 1 else 0
        from_l = torch.empty((n_from_l, 3), dtype=torch.float64, device=self.dev)
        from_r = torch.empty((n_from_r, 3), dtype=torch.float64, device=self.dev)
        ops = []
        if rank > 0:
            if cnt[rank, 0]:
                ops.append(dist.P2POp(dist.isend, send_l[:int(cnt[rank, 0])], rank - 1))
            if n_from_l:
                ops.append(dist.P2POp(dist.irecv, from_l, rank - 1))
        if rank < world - 1:
            if cnt[rank, 1]:
                ops.append(dist.P2POp(dist.isend, send_r[:int(cnt[rank, 1])], rank + 1))
            if n_from_r:
                ops.append(dist.P2POp(dist.irecv, from_r, rank + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # the tile's column lattice (the origin a single grid over the whole tile would use) and this strip's box:
        # its own points' box widened by the halo in x, the tile's range in y and z
        cell = cell_for_radius(self.radii[0])
        lattice = (float(gmin[0]) - cell, float(gmin[1]) - cell)
        bounds = (max(min(float(ai[rank, 0]), self.x_lo) - self.halo, float(gmin[0])), float(gmin[1]), float(gmin[2]),
                  min(max(float(ai[rank, 3]), self.x_hi) + self.halo, float(gmax[0])), float(gmax[1]), float(gmax[2]))
        step = StripStep(own, from_l if n_from_l else None, from_r if n_from_r else None, self.radii, self.params,
                         lattice=lattice, bounds=bounds)
        out = {}
        if out_host is None:
            allreduce_minmax_rows(step.features())           # ONE collective for all radii
            step.blend(self.cw)
            for r in self.radii:
                out[r] = step.saliency(r)
        else:
            # host buffers: one collective PER radius, so that every saliency column starts its way to the host as soon
            # as its radius is done -- the downloads spread over the step instead of piling up behind its end (with
            # eight ranks on one host a column takes 3-4 ms)
            step.build_grid()
            for i, r in enumerate(self.radii):
                step.features_radius(i)
                allreduce_minmax_rows(step.minmax[i:i + 1])
                step.blend_radius(i, self.cw)
                out[r] = step.saliency(r)
                _copy_out(out[r], out_host[r], copy_stream, self.dev)
        pc = step.pcount
        # the global Otsu threshold (Downsample.py:54) is NOT part of the saliency step (the single-GPU step has none
        # either): the step keeps its owned saliency columns and their local extrema, `otsu` runs the two collectives
        self._otsu_in = (dict(out), step.sal_minmax, step.radii)
        self._otsu = None
        res = _StepResult(step.radii, step.status, step.table_len, None, None)
        step.close()
        if self._last is not None:
            # the previous step's overflow words: read now, one step late, so that no step waits for its own
            need = self._last.overflow()
            if need:
                raise _lib.More3DError("dk/dn: chi2 table too short (%d needed)" % need, _lib.E_RANGE)
        self._last = res
        return out, pc

    @property
    def otsu(self):
        """{radius: global Otsu threshold of the saliency column} of the last step.  COLLECTIVE: every rank must read it
        (one all-reduce of the saliency extrema, 256-bin histograms over the global range, one all-reduce of the
        histograms, then the host scan)."""
        if self._otsu is None:
            lib = _lib.load()
            sal, smm, radii = self._otsu_in
            smm = smm.clone()
            allreduce_minmax_rows(smm)
            with torch.cuda.device(self.dev):
                hist = torch.empty((len(radii), 256), dtype=torch.int64, device=self.dev)
                for i, r in enumerate(radii):
                    v = sal[r].contiguous()
                    check(lib.more3d_histogram256_range(ptr(v), 4, int(v.numel()), ptr(smm[i]), 4, ptr(hist[i]), stream_ptr()))
            dist.all_reduce(hist, op=dist.ReduceOp.SUM)
            self._otsu = _StepResult(radii, None, 0, smm, hist).otsu()
        return self._otsu

    def step_resident(self):
        return self._run(self.resident, None)[0]

    def step_e2e(self):
        res = self._run(self.host.to(self.dev, non_blocking=True), self.out_host)[0]
        torch.cuda.current_stream().synchronize()
        return res

    def steps_e2e(self, k):
        """k strips one after the other from pinned host memory; the upload of strip i+1 runs on a copy stream
        while strip i is computed (the D2H of the saliency columns is issued per radius by the step)."""
        cur = torch.cuda.current_stream(self.dev)
        up = torch.cuda.Stream(device=self.dev)
        down = torch.cuda.Stream(device=self.dev)
        # two staging buffers filled alternately (no allocation between the strips, see pipeline.multiscale_saliency_tiles)
        if getattr(self, "_staging", None) is None:
            self._staging = [torch.empty_like(self.resident) for _ in range(2)]     # kept between calls
        bufs = self._staging[:min(2, max(k, 1))]
        freed = [None] * len(bufs)

        def upload(i):
            slot = i % len(bufs)
            with torch.cuda.stream(up):
                if freed[slot] is not None:
                    up.wait_event(freed[slot])
                else:
                    up.wait_stream(cur)
                bufs[slot].copy_(self.host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up)
            return bufs[slot], ev

        res = None
        nxt = upload(0)
        for i in range(k):
            d, ev = nxt
            if i + 1 < k:
                nxt = upload(i + 1)
            cur.wait_event(ev)
            res = self._run(d, self.out_host, copy_stream=down)[0]
            done = torch.cuda.Event()
            done.record(cur)
            freed[i % len(bufs)] = done
        cur.wait_stream(down)
        cur.synchronize()
        from .pipeline import release_copied
        release_copied()
        return res


# ------------------------------------------------------------------------------------------------
# voxel grid across ranks: voxel-aligned x-strips, owner = the voxel's strip, no halo
# ------------------------------------------------------------------------------------------------
def voxel_downsample_strips(comm, pts_own, voxel_size, gid_own=None):
    """open3d voxel_down_sample (Downsample.py:86-89) of a cloud that is spread over the ranks of `comm`
    (distributed_levelset.DistComm / ThreadComm), in any partition.

    One all-reduce makes the bounding box -- hence the voxel lattice -- global; the x range of the lattice is cut into
    `world` voxel-aligned strips; every point travels once to the rank that owns its voxel (for a cloud that is
    already in x-strips only the points next to a strip border move); each rank then reduces its own voxels with
    the single-GPU kernels on the global lattice.  No halo and no collective on the reduction itself: a voxel's
    members all sit on one rank.  Members are accumulated in the order of their global point ids (`gid_own`,
    default: rank-major numbering), so the union of the ranks' outputs equals the single-GPU output bit for bit.

    Returns (means (V, 3) f64, voxel indices (V, 3) int64) of this rank's voxels, device tensors."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    pts = torch.as_tensor(pts_own, dtype=torch.float64).to(dev).contiguous()
    n = int(pts.shape[0])
    world, rank = comm.world, comm.rank
    voxel = float(voxel_size)
    if not (voxel > 0 and np.isfinite(voxel)):
        raise ValueError("voxel_size must be > 0")
    # global bounding box: one MAX all-reduce over (max, -min)
    big = torch.finfo(torch.float64).max
    if n:
        ext = torch.cat([pts.amax(0), -pts.amin(0)])
    else:
        ext = torch.full((6,), -big, dtype=torch.float64, device=dev)
    comm.allreduce_max(ext)
    e = ext.cpu().numpy()
    gmax, gmin = e[:3], -e[3:]
    bounds = (C.c_double * 6)(*[float(v) for v in gmin], *[float(v) for v in gmax])
    vmin_x = float(gmin[0]) - voxel * 0.5                         # min_bound - voxel_size * 0.5 (as more3d_voxel_downsample)
    nx = int(np.floor((float(gmax[0]) - vmin_x) / voxel)) + 2
    # counts per (sender, receiver) so that everybody knows every message size: one SUM all-reduce
    if gid_own is None:
        cnt = torch.zeros(world, dtype=torch.float64, device=dev)
        cnt[rank] = n
        comm.allreduce_sum(cnt)
        start = int(cnt[:rank].sum().item())
        gid = torch.arange(start, start + n, dtype=torch.int64, device=dev)
    else:
        gid = torch.as_tensor(gid_own, dtype=torch.int64).to(dev)
    ix = torch.floor((pts[:, 0] - vmin_x) / voxel).to(torch.int64) if n else torch.zeros(0, dtype=torch.int64, device=dev)
    owner = torch.clamp((ix * world) // max(nx, 1), 0, world - 1)
    sel = {p: torch.nonzero(owner == p).flatten() for p in range(world)}
    m = torch.zeros((world, world), dtype=torch.float64, device=dev)
    for p in range(world):
        m[rank, p] = sel[p].numel()
    comm.allreduce_sum(m)
    mh = m.cpu().numpy().astype(np.int64)
    peers = [p for p in range(world) if p != rank]
    got_xyz = comm.exchange({p: pts[sel[p]] for p in peers}, {p: (int(mh[p, rank]), 3) for p in peers}, torch.float64, dev)
    got_gid = comm.exchange({p: gid[sel[p]] for p in peers}, {p: (int(mh[p, rank]),) for p in peers}, torch.int64, dev)
    mine = torch.cat([pts[sel[rank]]] + [got_xyz[p] for p in peers])
    mine_gid = torch.cat([gid[sel[rank]]] + [got_gid[p] for p in peers])
    nl = int(mine.shape[0])
    if nl == 0:
        return torch.zeros((0, 3), dtype=torch.float64, device=dev), torch.zeros((0, 3), dtype=torch.int64, device=dev)
    mine = mine[torch.argsort(mine_gid, stable=True)].contiguous()   # the whole cloud's point order
    with torch.cuda.device(dev):
        out = torch.empty((nl, 3), dtype=torch.float64, device=dev)
        keys = torch.empty((nl, 3), dtype=torch.int64, device=dev)
        nout = C.c_int64(0)
        check(lib.more3d_voxel_downsample_bounds(ptr(mine), nl, voxel, bounds, ptr(out), ptr(keys), None,
                                                 C.byref(nout), stream_ptr()))
    v = int(nout.value)
    return out[:v], keys[:v]
