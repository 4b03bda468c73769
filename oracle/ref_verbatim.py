"""The UNMODIFIED reference, executed in-process (test / bench infrastructure, NOT product code).

Loads the reference's own modules -- from the bytecode ``oracle/build_ref.py`` compiled into ``oracle/_ref/``
(what travels to the GPU box), else from the sources under /root/reference (build container) -- and drives
their per-point processors exactly as OPALS would: ``Curvature_Kernel.process(pt, neighbours)``
(DM_saliency.py:88-157) and ``Saliency_Kernel.process`` (:219-324) on duck-typed ``pt`` / ``neighbours``
objects fed from ``BallTreePointSet.queryRadius`` / ``sklearn BallTree.query_radius`` (self included).

Two in-process stubs make the modules importable without editing them: ``open3d`` (only
``PointCloud`` is touched, BallTreePointSet.py:32,37) and ``opals`` (``pyDM.KernelPointEx`` base class + empty
AddInfo / Info / Histo, DM_saliency.py:32-33).  The OPALS processor loop (DM_saliency.py:403-407) is replaced
by a plain Python loop.

Only ``tests/``, ``__graft_entry__.smoke()``, ``bench.py``'s CPU legs and ``tools/make_golden.py`` import this.
"""
import contextlib
import importlib.machinery
import importlib.util
import io
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILT = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("MORE3D_REFERENCE", "/root/reference")
_SRC_REL = {"DM_saliency": "DM_saliency.py", "levelsets_func": "levelsets_func.py",
            "BallTreePointSet": "Downsampling/BallTreePointSet.py"}
_MODS = {}


def available():
    """True when the verbatim reference can be loaded here (prebuilt bytecode or the source tree)."""
    return all(os.path.exists(os.path.join(REF_BUILT, m + ".bin")) for m in _SRC_REL) or \
        all(os.path.exists(os.path.join(REF_SRC, r)) for r in _SRC_REL.values())


def origin():
    """Where the reference modules are loaded from (goes into bench.py's `sample` string)."""
    if all(os.path.exists(os.path.join(REF_BUILT, m + ".bin")) for m in _SRC_REL):
        return "oracle/_ref bytecode of the unmodified reference"
    return "reference sources at " + REF_SRC


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


def _load(name):
    if name in _MODS:
        return _MODS[name]
    install_stubs()
    pyc = os.path.join(REF_BUILT, name + ".bin")
    src = os.path.join(REF_SRC, _SRC_REL[name])
    qual = "more3d_reference_" + name      # never shadows the product modules of the same name
    if os.path.exists(pyc):
        loader = importlib.machinery.SourcelessFileLoader(qual, pyc)
    elif os.path.exists(src):
        loader = importlib.machinery.SourceFileLoader(qual, src)
    else:
        raise ImportError("the verbatim reference is not available: run `python oracle/build_ref.py` in the build "
                          "container (neither %s nor %s exists)" % (pyc, src))
    spec = importlib.util.spec_from_loader(qual, loader)
    mod = importlib.util.module_from_spec(spec)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loader.exec_module(mod)
    _MODS[name] = mod
    return mod


def load_reference():
    """(BallTreePointSet module, DM_saliency module, levelsets_func module) of the reference."""
    bt_mod, dm_mod, ls_mod = _load("BallTreePointSet"), _load("DM_saliency"), _load("levelsets_func")
    ls_mod.verbose = False
    return bt_mod, dm_mod, ls_mod


class _Info:
    """Attribute row of one point; float_ columns round to float32 like OPALS ColumnType.float_
    (DM_saliency.py:381, 491-492)."""

    def __init__(self, store, i):
        self.s, self.i = store, i

    def get(self, c):
        return float(self.s[c][self.i])

    def set(self, c, v):
        self.s[c][self.i] = v  # numpy casts to the column dtype

    def isNull(self, c):
        return bool(np.isnan(self.s[c][self.i]))


class _Pt:
    def __init__(self, p, store, i):
        self.x, self.y, self.z = float(p[0]), float(p[1]), float(p[2])
        self._info = _Info(store, i)

    def info(self):
        return self._info


class _Nb:
    def __init__(self, d, m):
        self.d, self.m = d, m

    def sizePoint(self):
        return self.m

    def asNumpyDict(self):
        return self.d


def run_curvature(xyz, normals, nbrs, alpha, roughness, min_points, queries=None):
    """Verbatim Curvature_Kernel.process over the points `queries` (default: all).  `normals` has one row per
    query, `nbrs[t]` are the cloud indices of query t's neighbourhood.
    Returns (K f32, pcount i32, normals f64 -- flipped where the reference flips), one row per query."""
    dm_mod = _load("DM_saliency")
    q = np.arange(xyz.shape[0]) if queries is None else np.asarray(queries)
    n = len(q)
    nrm = np.array(normals, dtype=np.float64, copy=True)
    store = {0: nrm[:, 0], 1: nrm[:, 1], 2: nrm[:, 2],
             3: np.zeros(n, np.float32), 4: np.zeros(n, np.int32)}
    k = dm_mod.Curvature_Kernel()
    k.setArgs(alpha=alpha, roughness=roughness, min_points=min_points)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(n):
            j = nbrs[t]
            d = {"x": xyz[j, 0], "y": xyz[j, 1], "z": xyz[j, 2]}
            k.process(_Pt(xyz[q[t]], store, t), _Nb(d, len(j)))
    return store[3], store[4], nrm


def run_dk_dn(xyz, normals, K, nbrs, rho, sigma_nb, sigma_normal, sigma_curvature, alpha, min_points, queries=None):
    """Verbatim Saliency_Kernel.process over the points `queries` (default: all).  `normals` / `K` are the
    WHOLE cloud's columns (the neighbours' values are read from them).  Returns (dk f32, dn f32) per query."""
    dm_mod = _load("DM_saliency")
    q = np.arange(xyz.shape[0]) if queries is None else np.asarray(queries)
    n = len(q)
    normals = np.asarray(normals, dtype=np.float64)
    K = np.asarray(K, dtype=np.float32)
    nq = np.array(normals[q], dtype=np.float64, copy=True)
    store = {0: nq[:, 0], 1: nq[:, 1], 2: nq[:, 2], 3: K[q].copy(),
             4: np.zeros(n, np.float32), 5: np.zeros(n, np.float32)}
    s = dm_mod.Saliency_Kernel()
    s.setArgs(rho, sigma_nb, sigma_normal, sigma_curvature, alpha, min_points)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for t in range(n):
            j = nbrs[t]
            d = {"x": xyz[j, 0], "y": xyz[j, 1], "z": xyz[j, 2],
                 "_nonparam_K": K[j].astype(np.float64),
                 "NormalX": normals[j, 0], "NormalY": normals[j, 1], "NormalZ": normals[j, 2]}
            s.process(_Pt(xyz[q[t]], store, t), _Nb(d, len(j)))
    return store[4], store[5]
