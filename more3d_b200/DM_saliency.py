"""Host-side mirror of the reference's saliency module (DM_saliency.py) on top of the B200 kernels.

Same names, argument meaning and error behaviour as the reference:
``Curvature_Kernel`` / ``Saliency_Kernel`` (``setArgs`` as DM_saliency.py:59, :182),
``DM_nonparam_curvature`` (:326), ``DM_dk_dn`` (:419), ``DM_saliency`` (:539),
``DM_saliency_stats`` (:585).  Where the reference takes the path of an OPALS ODM file, these take
a :class:`PointCloudDM` (an in-memory, device-resident attribute table -- ODM is a proprietary
format), the path of an ``.npz`` written by ``PointCloudDM.save_as``, or an uncompressed ``.las`` file.

The OPALS per-point callback dispatch (``pyDM.ProcessorEx(dm, query, ...).run(kernel)``,
DM_saliency.py:403-407) is kept as :class:`ProcessorEx`; ``run`` launches ONE bulk CUDA pass that
evaluates ``kernel.process`` for every point instead of calling back into Python per point.
Only ``shape='sphere'`` is in scope (the reference marks circle/cylinder "TODO: check if ... works").
"""
import datetime
import re
import sys

import numpy as np
import torch
from scipy import stats

from . import _lib
from ._lib import check, ptr, stream_ptr
from .grid import GridColumn, UniformGrid, X_SPLIT, as_device_points, cell_for_radius


# Overflow status words of the dk/dn passes: handed out from one zeroed per-device ring (a word is only ever written on a
# chi2-table overflow, so the ring needs no re-zeroing kernel per pass; after an overflow the ring is replaced).
_STATUS_RING = {}
_RING_WORDS = 4096


def _status_words(device, count):
    key = str(device)
    ring = _STATUS_RING.get(key)
    if ring is None or ring[1] + count > _RING_WORDS:
        if ring is None:
            ring = [torch.zeros(_RING_WORDS, dtype=torch.int32, device=device), 0]
            _STATUS_RING[key] = ring
        else:
            ring[1] = 0                      # wrap: the oldest words were read long ago (and found zero)
    t = ring[0][ring[1]:ring[1] + count]
    ring[1] += count
    return t


_PINNED = [None, 0]


def _pinned_words(count):
    """`count` pinned int32 words from a ring that is allocated once (64 rows of 16; a row is read one tile after it was
    handed out, long before the ring comes round) -- a pinned allocation per tile costs a cudaHostAlloc now and then."""
    if count > 16:
        return torch.empty(count, dtype=torch.int32).pin_memory()
    if _PINNED[0] is None:
        _PINNED[0] = torch.zeros((64, 16), dtype=torch.int32).pin_memory()
    row = _PINNED[0][_PINNED[1] % 64]
    _PINNED[1] += 1
    return row[:count]


def _status_ring_poisoned(device):
    _STATUS_RING.pop(str(device), None)      # an overflow left a non-zero word: start a fresh ring


class _Columns(dict):
    """Attribute table: name -> device tensor in the caller's point order.  A value may be stored as a
    GridColumn (grid order); reading it through the mapping interface un-permutes it once, `raw` does not."""

    def raw(self, name):
        return dict.get(self, name)

    def __getitem__(self, name):
        v = dict.__getitem__(self, name)
        return v.tensor() if isinstance(v, GridColumn) else v

    def get(self, name, default=None):
        return self[name] if name in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class PointCloudDM:
    """In-memory stand-in for ``pyDM.Datamanager``: xyz plus named per-point attribute columns,
    resident in HBM.  Column names follow the reference layout (DM_saliency.py:377-384, 486-494)."""

    FLOAT_COLS = ("_nonparam_K", "_dk", "_dn", "_normed_dk", "_normed_dn", "_saliency")

    def __init__(self, xyz, normals=None, device=None):
        self.xyz = as_device_points(xyz, device)
        self.device = self.xyz.device
        self.n = int(self.xyz.shape[0])
        self.attrs = _Columns()
        self._normals = None if normals is None else as_device_points(normals, self.device).clone()
        self._grids = {}
        self._minmax = None
        # chi2-table overflow words of dk/dn passes that have not been checked yet (defer_checks: the pipeline reads
        # them once at its end instead of synchronising after every radius)
        self.defer_checks = False
        self._pending = []
        self._status = None

    @property
    def normals(self):
        """(N,3) f64 device tensor in point order (un-permuted on first access if held in grid order)."""
        v = self._normals
        return v.tensor() if isinstance(v, GridColumn) else v

    @normals.setter
    def normals(self, value):
        self._normals = value

    # -- ODM-like surface ---------------------------------------------------------------
    @classmethod
    def load(cls, path, *_):
        if str(path).lower().endswith(".las"):
            # an uncompressed LAS file: coordinates as stored (not centred), NormalX/Y/Z and every other
            # extra-bytes dimension as an attribute column
            from .io_formats import parse_las
            try:
                las = parse_las(path)
            except (OSError, ValueError) as e:
                raise IOError(str(e))
            nrm = None
            if all(k in las.extra for k in ("NormalX", "NormalY", "NormalZ")):
                nrm = np.column_stack([las.extra[k] for k in ("NormalX", "NormalY", "NormalZ")]).astype(np.float64)
            dm = cls(np.column_stack((las.x, las.y, las.z)), nrm)
            for k, v in las.extra.items():
                if k not in ("NormalX", "NormalY", "NormalZ"):
                    dm.attrs[k] = torch.from_numpy(np.ascontiguousarray(v)).to(dm.device)
            return dm
        try:
            z = np.load(path)
        except (OSError, ValueError) as e:
            raise IOError(str(e))
        dm = cls(z["xyz"], z["normals"] if "normals" in z else None)
        for k in z.files:
            if k not in ("xyz", "normals"):
                dm.attrs[k] = torch.from_numpy(z[k]).to(dm.device)
        return dm

    def save(self):
        """ODM persistence is a no-op for the in-memory table (DM_saliency.py:415, :527)."""
        return self

    def save_as(self, path):
        out = {"xyz": self.xyz.cpu().numpy()}
        if self._normals is not None:
            out["normals"] = self.normals.cpu().numpy()
        for k, v in self.attrs.items():
            out[k] = v.cpu().numpy()
        np.savez(path, **out)

    def sizePoint(self):
        return self.n

    def get(self, name):
        """Column as a host numpy array."""
        if name in ("NormalX", "NormalY", "NormalZ"):
            return self.normals[:, "XYZ".index(name[-1])].cpu().numpy()
        return self.attrs[name].cpu().numpy()

    # -- spatial index ------------------------------------------------------------------
    def grid(self, radius):
        """Uniform grid for fixed-radius work at `radius`.  A grid built for a smaller radius is reused as
        long as its columns are at most 2x finer than this radius would ask for (a 9x9 stencil of quarter-
        radius columns costs the same as 5x5 half-radius ones -- measured -- and saves the rebuild)."""
        want = cell_for_radius(radius)
        best = None
        for g in self._grids.values():
            # keyed on the REQUESTED cell: on a wide tile the library widens the columns (g.cell > requested), such a
            # grid still serves every radius it was asked for -- the stencil half-width follows g.cell
            c = g.requested_cell
            if c <= want * (1 + 1e-12) and want / c <= 2.05 and (best is None or c > best.requested_cell):
                best = g
        if best is None:
            best = UniformGrid(self.xyz, want, x_split=X_SPLIT)
            self._grids[float(radius)] = best
        return best

    def status_word(self):
        """A zeroed device int32 for the next dk/dn pass's overflow report (slot len(_pending) of one small tensor)."""
        if self._status is None or len(self._pending) >= int(self._status.numel()):
            self._status = _status_words(self.device, 16)
            self._pending = []
        i = len(self._pending)
        return self._status[i:i + 1]

    def _need(self, host_words):
        need = 0
        for v, length in zip(host_words, self._pending):
            if int(v) != 0:
                _status_ring_poisoned(self.device)
            if int(v) > length:
                need = max(need, int(v))
        self._pending = []
        self._status = None
        return need

    def check_pending(self):
        """Read the deferred overflow words (one synchronisation).  Returns 0 if every dk/dn pass since the last
        check had a long enough chi2 table, else the table length that would have sufficed."""
        if not self._pending:
            return 0
        return self._need(self._status[:len(self._pending)].cpu().tolist())

    def check_pending_async(self, side_stream):
        """Like check_pending, but the words travel on `side_stream` behind everything queued so far on the current
        stream; returns a callable that waits for that copy only (not for work queued later) and gives the answer."""
        if not self._pending:
            return lambda: 0
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        words = self._status[:len(self._pending)]
        host = _pinned_words(len(self._pending))
        with torch.cuda.stream(side_stream):
            side_stream.wait_event(ev)
            host.copy_(words, non_blocking=True)
            words.record_stream(side_stream)
            done = torch.cuda.Event()
            done.record(side_stream)

        # the closure must not keep `self` (and with it every column of this cloud) alive until it is called: it gets the
        # few values it needs, the cloud's own bookkeeping is reset now
        pending, device = list(self._pending), self.device
        self._pending = []
        self._status = None

        def result():
            done.synchronize()
            need = 0
            for v, length in zip(host.tolist(), pending):
                if int(v) != 0:
                    _status_ring_poisoned(device)
                if int(v) > length:
                    need = max(need, int(v))
            return need
        return result

    def drop_grids(self):
        for g in self._grids.values():
            g.close()
        self._grids = {}


class QueryDescriptor:
    """``pyDM.QueryDescriptor('sphere(r=0.5)')`` (DM_saliency.py:393-398)."""

    def __init__(self, shape):
        m = re.fullmatch(r"\s*sphere\(r=([^)]+)\)\s*", str(shape))
        if not m:
            raise NotImplementedError("only 'sphere(r=...)' queries are in scope, got %r" % (shape,))
        self.radius = float(m.group(1))


class ProcessorEx:
    """``pyDM.ProcessorEx(dm, query, processFilter, processLayout, processLayoutReadOnly,
    spatialFilter, spatialLayout, spatialLayoutReadOnly)`` (DM_saliency.py:403, :513-516)."""

    def __init__(self, dm, query, processFilter=None, processLayout=None, processLayoutReadOnly=False,
                 spatialFilter=None, spatialLayout=None, spatialLayoutReadOnly=False):
        if processFilter is not None or spatialFilter is not None:
            raise NotImplementedError("process/spatial filters are out of scope")
        self.dm, self.query = dm, query

    def run(self, kernel):
        kernel.tileFilter()
        kernel.process_all(self.dm, self.query.radius)
        kernel.releaseLeaf()


class _KernelPointEx:
    def tileFilter(self):
        return None

    def releaseLeaf(self):
        return

    def leafChanged(self, leaf, *args, **kwargs):
        return


class Curvature_Kernel(_KernelPointEx):
    """Non-parametric curvature processor (DM_saliency.py:40-157)."""
    _roughness = None
    _alpha = None
    _minPoints_neighbourhood = None

    def setArgs(self, alpha, roughness, min_points=4):
        self._alpha = alpha
        self._roughness = roughness
        self._minPoints_neighbourhood = min_points

    def epsilon(self):
        return float(stats.norm.ppf(1 - self._alpha / 2) * self._roughness)   # DM_saliency.py:99

    def process_all(self, dm, radius):
        """``process(pt, neighbours)`` for every point of ``dm`` in one device pass.
        Uses the normals stored in ``dm``; if there are none (the reference requires them,
        DM_saliency.py:329-330) centroid-PCA normals at the same radius are computed in the same pass."""
        g = dm.grid(radius)
        if dm._normals is None:
            # outputs stay in grid order (GridColumn): coalesced writes, un-permuted only if somebody reads them
            normals, ratio, K, pcount = g.normals_curvature_gridorder(radius, self.epsilon(),
                                                                      self._minPoints_neighbourhood)
            dm.normals = normals
            dm.attrs["_lam_ratio"] = ratio
        else:
            K, pcount = g.nonparam_curvature(radius, dm.normals, self.epsilon(), self._minPoints_neighbourhood)
        dm.attrs["_nonparam_K"] = K
        dm.attrs["_pcount"] = pcount
        return True


_CHI2_CACHE = {}


def _chi2_table(alpha, length, device):
    key = (float(alpha), int(length), str(device))
    t = _CHI2_CACHE.get(key)
    if t is None:
        with np.errstate(all="ignore"):
            host = stats.chi2.ppf(alpha, np.arange(length)).astype(np.float64)   # DM_saliency.py:282
        t = torch.from_numpy(host).to(device)
        _CHI2_CACHE[key] = t
    return t


class Saliency_Kernel(_KernelPointEx):
    """dk / dn processor (DM_saliency.py:159-324)."""
    _alpha = None
    _sigma_curvature = None
    _sigma_normal = None
    _sigma_neighbourhood = None
    _rho_neighbourhood = None
    _minPoints_neighbourhood = None

    def setArgs(self, rho_neighbourhood, sigma_neighbourhood, sigma_normal=0.01, sigma_curvature=0.1,
                alpha=0.05, min_points=4):
        self._alpha = alpha
        self._sigma_curvature = sigma_curvature
        self._sigma_normal = sigma_normal
        self._sigma_neighbourhood = sigma_neighbourhood   # stored, unused -- as in the reference (:268-269)
        self._rho_neighbourhood = rho_neighbourhood
        self._minPoints_neighbourhood = min_points

    def process_all(self, dm, radius):
        if dm._normals is None or "_nonparam_K" not in dm.attrs:
            raise ValueError("dk/dn needs normals and _nonparam_K (run DM_nonparam_curvature first)")
        g = dm.grid(radius)
        rn, rk = dm._normals, dm.attrs.raw("_nonparam_K")
        # the columns are still the untouched grid-order outputs of the fused pass on THIS grid: its cached
        # feature records are current, and dk/dn can stay in grid order too
        cached = isinstance(rn, GridColumn) and isinstance(rk, GridColumn) and rn.grid is g and rk.grid is g \
            and getattr(g, "_feat_cols", None) == (rn, rk) and rn.clean() and rk.clean()
        length = int(getattr(dm, "chi2_table_len", 1024))
        while True:
            table = _chi2_table(self._alpha, length, dm.device)
            status = dm.status_word()
            if cached:
                dk, dn, minmax = g.dk_dn_gridorder(radius, self._rho_neighbourhood, self._sigma_normal,
                                                   self._sigma_curvature, table, self._minPoints_neighbourhood,
                                                   status=status)
            else:
                dk, dn, minmax = g.dk_dn(radius, dm.normals, dm.attrs["_nonparam_K"], self._rho_neighbourhood,
                                         self._sigma_normal, self._sigma_curvature, table,
                                         self._minPoints_neighbourhood, status=status)
            if dm.defer_checks:
                dm._pending.append(length)                       # read by PointCloudDM.check_pending()
                break
            need = int(status.item())         # a neighbourhood with more active neighbours than the table covers
            dm._pending, dm._status = [], None
            if need:
                _status_ring_poisoned(dm.device)
            if need <= length:
                break
            if need > (1 << 24):
                raise _lib.More3DError("dk/dn: %d active neighbours in one neighbourhood" % (need - 1), _lib.E_RANGE)
            length = max(8 * length, need)
        dm.attrs["_dk"] = dk
        dm.attrs["_dn"] = dn
        # the fused extrema are valid for exactly these two column objects in their present state
        ver = lambda c: c.data._version if isinstance(c, GridColumn) else c._version
        dm._minmax = (minmax, dk, dn, ver(dk), ver(dn))
        return True


def _open(inFile):
    if isinstance(inFile, PointCloudDM):
        return inFile
    try:
        return PointCloudDM.load(inFile, False, False)
    except IOError as e:
        print(e)
        return None


def _check_shape(shape, kwargs):
    # sanity checks of DM_saliency.py:367-374 / :476-483
    if "radius" not in kwargs:
        print("must have radius for search")
        sys.exit(1)
    if shape == "cylinder":
        if "zmin" not in kwargs or "zmax" not in kwargs:
            print("for cylinder zmin and zmax must be defined for search")
            sys.exit(1)
    if shape != "sphere":
        raise NotImplementedError("only shape='sphere' is in scope (reference: circle/cylinder are TODO)")
    return "sphere(r=" + str(kwargs["radius"]) + ")"


def DM_nonparam_curvature(inFile, shape, roughness=0.15, alpha=0.05, min_points=3, verbose=False, **kwargs):
    """DM_saliency.py:326-417.  Returns the data manager."""
    dm = _open(inFile)
    if dm is None:
        return
    query = QueryDescriptor(_check_shape(shape, kwargs))
    k = Curvature_Kernel()
    k.setArgs(alpha=alpha, roughness=roughness, min_points=min_points)
    start = datetime.datetime.now()
    processor = ProcessorEx(dm, query, None, None, False, None, None, False)
    processor.run(k)
    if verbose:
        torch.cuda.synchronize(dm.device)
        print("Done. Non-parametric curvature computation took %.2f [s]"
              % (datetime.datetime.now() - start).total_seconds())
    dm.save()
    return dm


def DM_dk_dn(inFile, shape, rho_neighbourhood, sigma_neighbourhood, sigma_normal=0.01, sigma_curvature=0.1,
             alpha=0.05, min_points=4, verbose=False, **kwargs):
    """DM_saliency.py:419-527."""
    dm = _open(inFile)
    if dm is None:
        return
    query = QueryDescriptor(_check_shape(shape, kwargs))
    s = Saliency_Kernel()
    s.setArgs(rho_neighbourhood, sigma_neighbourhood, sigma_normal, sigma_curvature, alpha, min_points)
    processor = ProcessorEx(dm, query, None, None, False, None, None, False)
    start = datetime.datetime.now()
    processor.run(s)
    if verbose:
        torch.cuda.synchronize(dm.device)
        print("Done. dn and dk computation took %.2f [s]" % (datetime.datetime.now() - start).total_seconds())
    dm.save()
    return dm


def DM_saliency(inFile, curvature_weight=.5, verbose=False, **kwargs):
    """DM_saliency.py:539-583: Histo min/max of _dk/_dn, the two normalisations and the blend."""
    dm = _open(inFile)
    if dm is None:
        return
    lib = _lib.load()
    rdk, rdn = dm.attrs.raw("_dk"), dm.attrs.raw("_dn")
    # normalise + blend are elementwise: when both inputs are untouched grid-order columns of one grid the blend
    # runs in grid order and only _saliency is ever brought back to point order
    grid = rdk if (isinstance(rdk, GridColumn) and isinstance(rdn, GridColumn) and rdk.order is rdn.order
                   and rdk.clean() and rdn.clean()) else None
    dk, dn = (rdk.data, rdn.data) if grid is not None else (dm.attrs["_dk"], dm.attrs["_dn"])
    with torch.cuda.device(dm.device):
        minmax = None
        mm = dm._minmax
        if mm is not None and mm[1] is rdk and mm[2] is rdn:
            # reuse the extrema the dk/dn pass reduced, unless a column was edited in place since (the reference
            # recomputes them from the stored columns on every call, DM_saliency.py:562-568)
            ver = lambda c: c.data._version if isinstance(c, GridColumn) else c._version
            fresh = ver(rdk) == mm[3] and ver(rdn) == mm[4]
            if fresh and isinstance(rdk, GridColumn):
                fresh = rdk.clean() and rdn.clean()
            if fresh:
                minmax = mm[0]
        if minmax is None:
            minmax = torch.empty(4, dtype=torch.float32, device=dm.device)
            check(lib.more3d_minmax4(ptr(dk), ptr(dn), dm.n, ptr(minmax), stream_ptr()))
        sal = torch.empty(dm.n, dtype=torch.float32, device=dm.device)
        smm = torch.empty(2, dtype=torch.float32, device=dm.device)
        check(lib.more3d_saliency_blend(ptr(dk), ptr(dn), dm.n, ptr(minmax), float(curvature_weight),
                                        ptr(sal), ptr(smm), stream_ptr()))
    dm.attrs["_saliency"] = sal if grid is None else GridColumn(grid, sal)
    dm.attrs["_saliency_minmax"] = smm
    dm.attrs["_dkdn_minmax"] = minmax
    if verbose:
        mm = minmax.cpu().numpy()
        print("diff_k {}, diff_n {}".format(float(mm[1]) - float(mm[0]), float(mm[3]) - float(mm[2])))
    return dm


def DM_saliency_stats(odm):
    """DM_saliency.py:585-607: (min, max, mean, sigma) of ``_saliency``."""
    dm = _open(odm)
    if dm is None:
        print("Unable to open ODM '" + str(odm) + "'")
        sys.exit(1)
    if "_saliency" not in dm.attrs:
        print("no saliency values")
        return 1
    sal = dm.attrs["_saliency"].contiguous()
    with torch.cuda.device(sal.device):
        out = torch.empty(4, dtype=torch.float64, device=sal.device)
        check(_lib.load().more3d_column_stats(ptr(sal), int(sal.numel()), ptr(out), stream_ptr()))
    return tuple(float(v) for v in out.cpu().numpy())
