"""CPU: the C-ABI library loads and exports every symbol include/more3d_b200.h declares; host-side
argument handling that needs no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "more3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(more3d_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("more3d_grid_create", "more3d_radius_count", "more3d_radius_fill", "more3d_knn",
                 "more3d_normals_pca", "more3d_nonparam_curvature", "more3d_dk_dn", "more3d_saliency_blend"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from more3d_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.more3d_version() >= 100
    # every declared int-returning entry point has a ctypes signature in the binding
    unbound = [n for n in declared_symbols()
               if n not in _lib.SIGNATURES and n not in ("more3d_version", "more3d_last_error", "more3d_launch_count",
                                                        "more3d_ls_scratch_doubles")]
    assert not unbound, unbound


def test_argument_errors_need_no_device():
    from more3d_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.more3d_grid_create(None, 0, 1.0, None, ctypes.byref(h))
    assert rc == -1 and b"empty" in lib.more3d_last_error()
    with pytest.raises(_lib.More3DError):
        _lib.check(rc)
    assert lib.more3d_grid_destroy(None) == 0


def test_host_mirror_error_behaviour(capsys):
    from more3d_b200 import DM_saliency as dm
    # unreadable ODM path: print the IOError and return None (reference DM_saliency.py:361-365)
    assert dm.DM_nonparam_curvature("/nonexistent/file.npz", "sphere", radius=0.5) is None
    assert dm.DM_dk_dn("/nonexistent/file.npz", "sphere", 0.3, 0.3, radius=0.5) is None
    assert "nonexistent" in capsys.readouterr().out
    q = dm.QueryDescriptor("sphere(r=0.75)")
    assert q.radius == 0.75
    with pytest.raises(NotImplementedError):
        dm.QueryDescriptor("circle(r=1)")
    with pytest.raises(SystemExit):
        dm._check_shape("sphere", {})                  # "must have radius for search", :368-370
    with pytest.raises(SystemExit):
        dm._check_shape("cylinder", {"radius": 1.0})   # :371-374
    k = dm.Curvature_Kernel()
    k.setArgs(alpha=0.05, roughness=0.02)
    assert k._minPoints_neighbourhood == 4 and abs(k.epsilon() - 0.0391992796908) < 1e-12
    s = dm.Saliency_Kernel()
    s.setArgs(0.3, 0.3)
    assert (s._sigma_normal, s._sigma_curvature, s._alpha, s._minPoints_neighbourhood) == (0.01, 0.1, 0.05, 4)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from more3d_b200 import _lib
    from more3d_b200.BallTreePointSet import BallTreePointSet
    with pytest.raises(_lib.More3DError):
        BallTreePointSet(np.zeros((10, 3)))
    with pytest.raises(ValueError):
        BallTreePointSet([[0, 0, 0]])                  # "Wrong input object.", BallTreePointSet.py:40-42


def header_prototypes():
    """name -> list of parameter type strings, parsed from the header"""
    src = open(os.path.join(ROOT, "include", "more3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\s*\*)\s+(more3d_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[m.group(1)] = params
    return protos


def test_ctypes_signatures_match_the_header():
    """Every binding in more3d_b200/_lib.py passes as many arguments as the header declares, and passes a
    pointer / integer / double where the header has one (a mismatch would corrupt the call silently)."""
    from more3d_b200 import _lib
    protos = header_prototypes()
    assert len(protos) >= 50
    bad = []
    for name, argtypes in _lib.SIGNATURES.items():
        params = protos.get(name)
        if params is None:
            bad.append((name, "not in header"))
            continue
        if len(params) != len(argtypes):
            bad.append((name, "header %d args, binding %d" % (len(params), len(argtypes))))
            continue
        for i, (p, t) in enumerate(zip(params, argtypes)):
            is_ptr = "*" in p
            if is_ptr != (t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "contents")):
                bad.append((name, "arg %d: header %r vs %r" % (i, p, t)))
            elif not is_ptr and re.match(r"(const\s+)?double\b", p) and t is not ctypes.c_double:
                bad.append((name, "arg %d: header %r vs %r" % (i, p, t)))
            elif not is_ptr and re.match(r"(const\s+)?int64_t\b", p) and t is not ctypes.c_int64:
                bad.append((name, "arg %d: header %r vs %r" % (i, p, t)))
            elif not is_ptr and re.match(r"(const\s+)?int\b", p) and t is not ctypes.c_int:
                bad.append((name, "arg %d: header %r vs %r" % (i, p, t)))
    assert not bad, bad


def test_ls_state_struct_matches_the_ctypes_mirror():
    """more3d_ls_state (header) and distributed_levelset._State (ctypes) list the same fields in the same order."""
    from more3d_b200.distributed_levelset import _State
    src = open(os.path.join(ROOT, "include", "more3d_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} more3d_ls_state;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names, kinds = [], []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ptr = "*" in decl
        base = re.match(r"(const\s+)?(\w+)", decl).group(2)
        for n in re.split(r",", decl[decl.index(base) + len(base):]):
            names.append(n.replace("*", "").strip())
            kinds.append("ptr" if ptr else base)
    got = [(n, "ptr" if t is ctypes.c_void_p else {ctypes.c_int64: "int64_t", ctypes.c_double: "double",
                                                     ctypes.c_int: "int"}[t]) for n, t in _State._fields_]
    assert got == list(zip(names, kinds))


def test_screening_thresholds_reproduce_the_reference_active_test():
    """The screening pass of dk/dn replaces `w > 0.01` (DM_saliency.py:272-279) by two thresholds on the squared distance
    (more3d_dkdn_act_interval, host only).  Against the reference's own expression in numpy float64: each threshold sits
    exactly at the edge of its branch, and random distances agree with the expression on both sides."""
    import numpy as np
    from more3d_b200 import _lib
    lib = _lib.load()
    rng = np.random.Generator(np.random.PCG64(5))
    for rho in [0.5 / np.sqrt(2), 0.75 / np.sqrt(2), 1.0 / np.sqrt(2), 0.3, 1.234567, 1e-3, 37.5]:
        lo, hi = ctypes.c_double(0), ctypes.c_double(0)
        _lib.check(lib.more3d_dkdn_act_interval(float(rho), ctypes.byref(lo), ctypes.byref(hi)))
        lo, hi = lo.value, hi.value
        rho2 = np.float64(rho) ** 2

        def active(d):                           # the reference's expression, element by element
            d = np.asarray(d, np.float64)
            w = np.zeros_like(d)
            a, b = d < rho2, d > rho2
            w[a] = 1 / rho2 * d[a]
            w[b] = -1 / rho2 * d[b] + 2
            w[w < 0] = 0
            return w > 0.01

        assert lo < rho2 < hi
        assert active([lo])[0] and not active([np.nextafter(lo, 0.0)])[0]
        assert active([hi])[0] and not active([np.nextafter(hi, np.inf)])[0]
        d = np.concatenate([rng.random(20000) * 2.2 * rho2, lo * (1 + (rng.random(2000) - 0.5) * 1e-13),
                            hi * (1 + (rng.random(2000) - 0.5) * 1e-13), [rho2, 0.0, 2 * rho2]])
        assert np.array_equal(active(d), (d >= lo) & (d <= hi) & (d != rho2))
    assert lib.more3d_dkdn_act_interval(0.0, ctypes.byref(ctypes.c_double()), ctypes.byref(ctypes.c_double())) == -1
