"""Run the UNMODIFIED reference (/root/reference) in-process to produce golden vectors.

Only usable in the build container (the GPU box has no /root/reference).  Nothing under
``tests/``, ``bench.py`` or the product imports this at run time; ``tools/make_golden.py`` uses it
to write ``tests/golden/*.npz``.

Two in-process stubs make the reference modules importable without editing them
(SURVEY.md §8c): ``open3d`` (only ``PointCloud`` is touched, BallTreePointSet.py:32,37) and
``opals`` (``pyDM.KernelPointEx`` base class + empty AddInfo/Info/Histo, DM_saliency.py:32-33).
The OPALS processor loop (DM_saliency.py:403-407) is replaced by a plain Python loop that feeds
``process(pt, neighbours)`` from ``BallTreePointSet.queryRadius`` (self included).
"""
import sys
import types
import io
import contextlib
import warnings

import numpy as np

REF = "/root/reference"


def install_stubs():
    if "open3d" not in sys.modules:
        o3d = types.ModuleType("open3d")

        class PointCloud:  # never instantiated; isinstance() target only
            pass

        o3d.PointCloud = PointCloud
        sys.modules["open3d"] = o3d
    if "opals" not in sys.modules:
        opals = types.ModuleType("opals")
        pyDM = types.ModuleType("opals.pyDM")

        class KernelPointEx:
            def __init__(self):
                pass

        pyDM.KernelPointEx = KernelPointEx
        opals.pyDM = pyDM
        sys.modules["opals"] = opals
        sys.modules["opals.pyDM"] = pyDM
        for name in ("AddInfo", "Info", "Histo"):
            m = types.ModuleType("opals." + name)
            setattr(opals, name, m)
            sys.modules["opals." + name] = m
    for p in (REF, REF + "/Downsampling"):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_reference():
    install_stubs()
    import BallTreePointSet as bt_mod  # noqa
    import DM_saliency as dm_mod  # noqa
    import levelsets_func as ls_mod  # noqa
    ls_mod.verbose = False
    return bt_mod, dm_mod, ls_mod


class _Info:
    """Attribute row of one point; float_ columns round to float32 like OPALS ColumnType.float_
    (DM_saliency.py:381, 491-492)."""

    def __init__(self, store, i, f32_cols):
        self.s, self.i, self.f32 = store, i, f32_cols

    def get(self, c):
        return float(self.s[c][self.i])

    def set(self, c, v):
        self.s[c][self.i] = v  # numpy casts to the column dtype

    def isNull(self, c):
        return bool(np.isnan(self.s[c][self.i]))


class _Pt:
    def __init__(self, xyz, store, i, f32_cols):
        self.x, self.y, self.z = float(xyz[i, 0]), float(xyz[i, 1]), float(xyz[i, 2])
        self._info = _Info(store, i, f32_cols)

    def info(self):
        return self._info


class _Nb:
    def __init__(self, d, m):
        self.d, self.m = d, m

    def sizePoint(self):
        return self.m

    def asNumpyDict(self):
        return self.d


def run_curvature(xyz, normals, nbrs, alpha, roughness, min_points):
    """Verbatim Curvature_Kernel.process over every point.  Returns (K f32, pcount i32, normals f64)."""
    _, dm_mod, _ = load_reference()
    n = xyz.shape[0]
    nrm = np.array(normals, dtype=np.float64, copy=True)
    store = {0: nrm[:, 0], 1: nrm[:, 1], 2: nrm[:, 2],
             3: np.zeros(n, np.float32), 4: np.zeros(n, np.int32)}
    k = dm_mod.Curvature_Kernel()
    k.setArgs(alpha=alpha, roughness=roughness, min_points=min_points)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            j = nbrs[i]
            d = {"x": xyz[j, 0], "y": xyz[j, 1], "z": xyz[j, 2]}
            k.process(_Pt(xyz, store, i, (3,)), _Nb(d, len(j)))
    return store[3], store[4], nrm


def run_dk_dn(xyz, normals, K, nbrs, rho, sigma_nb, sigma_normal, sigma_curvature, alpha, min_points):
    """Verbatim Saliency_Kernel.process over every point.  Returns (dk f32, dn f32)."""
    _, dm_mod, _ = load_reference()
    n = xyz.shape[0]
    nrm = np.array(normals, dtype=np.float64, copy=True)
    K = np.asarray(K, dtype=np.float32)
    store = {0: nrm[:, 0], 1: nrm[:, 1], 2: nrm[:, 2], 3: K.copy(),
             4: np.zeros(n, np.float32), 5: np.zeros(n, np.float32)}
    s = dm_mod.Saliency_Kernel()
    s.setArgs(rho, sigma_nb, sigma_normal, sigma_curvature, alpha, min_points)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for i in range(n):
            j = nbrs[i]
            d = {"x": xyz[j, 0], "y": xyz[j, 1], "z": xyz[j, 2],
                 "_nonparam_K": K[j].astype(np.float64),
                 "NormalX": normals[j, 0], "NormalY": normals[j, 1], "NormalZ": normals[j, 2]}
            s.process(_Pt(xyz, store, i, (3, 4, 5)), _Nb(d, len(j)))
    return store[4], store[5]
