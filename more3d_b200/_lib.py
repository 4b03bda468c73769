"""ctypes binding of libmore3d_b200.so (the C-ABI declared in include/more3d_b200.h).

There is NO CPU fallback: if the shared library is missing and cannot be built, or no CUDA
device is present, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmore3d_b200.so")

_lib = None

_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_double

# name -> argtypes; every function returns int except the two noted below
SIGNATURES = {
    "more3d_profile_enable": [_i32],
    "more3d_profile_reset": [],
    "more3d_profile_get": [C.c_char_p, C.POINTER(_f64), C.POINTER(_i64)],
    "more3d_profile_names": [C.c_char_p, _i32],
    "more3d_read_back": [_vp, _vp, _i64, _vp],
    "more3d_set_dkdn_variant": [_i32],
    "more3d_set_dkdn_gate": [_f64],
    "more3d_dkdn_act_interval": [_f64, C.POINTER(_f64), C.POINTER(_f64)],
    "more3d_grid_create": [_vp, _i64, _f64, _vp, C.POINTER(_vp)],
    "more3d_grid_create_segments": [C.POINTER(_vp), C.POINTER(_i64), _i32, _f64, _i32, C.POINTER(_f64), C.POINTER(_f64),
                                    _vp, C.POINTER(_vp)],
    "more3d_grid_destroy": [_vp],
    "more3d_grid_info": [_vp, C.POINTER(_i64), C.POINTER(_f64)],
    "more3d_grid_order": [_vp, _vp, _vp],
    "more3d_radius_count": [_vp, _vp, _i64, _f64, _vp, _vp],
    "more3d_offsets_from_counts": [_vp, _i64, _vp, C.POINTER(_i64), _vp],
    "more3d_radius_fill": [_vp, _vp, _i64, _f64, _vp, _vp, _i32, _vp],
    "more3d_knn": [_vp, _vp, _i64, _i32, _vp, _vp, _vp],
    "more3d_normals_pca": [_vp, _f64, _vp, _vp, _vp, _vp],
    "more3d_nonparam_curvature": [_vp, _f64, _vp, _f64, _i32, _vp, _vp, _vp],
    "more3d_normals_curvature": [_vp, _f64, _f64, _i32, _vp, _vp, _vp, _vp, _vp],
    "more3d_dk_dn": [_vp, _f64, _vp, _vp, _f64, _f64, _f64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "more3d_normals_curvature_gridorder": [_vp, _f64, _f64, _i32, _vp, _vp, _vp, _vp, _vp],
    "more3d_dk_dn_gridorder": [_vp, _f64, _f64, _f64, _f64, _vp, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp],
    "more3d_unpermute": [_vp, _i64, _vp, _vp, _i32, _vp],
    "more3d_minmax4": [_vp, _vp, _i64, _vp, _vp],
    "more3d_saliency_blend": [_vp, _vp, _i64, _vp, _f64, _vp, _vp, _vp],
    "more3d_histogram256": [_vp, _i32, _i64, _vp, _vp, _vp],
    "more3d_voxel_downsample": [_vp, _i64, _f64, _vp, _vp, _vp, C.POINTER(_i64), _vp],
    "more3d_voxel_downsample_bounds": [_vp, _i64, _f64, C.POINTER(_f64), _vp, _vp, _vp, C.POINTER(_i64), _vp],
    "more3d_balltree_shape": [_i64, _i32, C.POINTER(_i64), C.POINTER(_i64)],
    "more3d_balltree_build": [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "more3d_lod_select": [_vp, _vp, _vp, _i64, _i64, _f64, _vp, _vp, C.POINTER(_i64), _vp],
    "more3d_lod_cell_rule": [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _i32, _vp, _vp, _vp, _vp],
    "more3d_lod_assemble": [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i64), _vp],
    "more3d_select_flagged": [_vp, _i64, _vp, C.POINTER(_i64), _vp],
    "more3d_ls_frames": [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp],
    "more3d_ls_graph_create": [_vp, _i64, _i32, _vp, _vp, _vp, _f64, _i32, _vp, _vp, C.POINTER(_i32), _vp, C.POINTER(_vp)],
    "more3d_ls_graph_destroy": [_vp],
    "more3d_ls_diff": [_vp, _vp, _vp, _vp, _vp],
    "more3d_ls_smooth": [_vp, _vp, _vp, _vp, _vp],
    "more3d_ls_zero_crossings": [_vp, _vp, _vp, _vp, _vp],
    "more3d_ls_init_voxels": [_vp, _i64, _f64, _f64, _vp, _vp],
    "more3d_ls_run": [_vp, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _f64, _f64, _f64, _f64, _f64, _f64, _f64, _i32,
                      _f64, C.POINTER(_f64), C.POINTER(_i32), C.POINTER(_i32), _vp],
    "more3d_ls_extract": [_vp, _vp, _vp, _i32, _i32, _vp, _vp],
    "more3d_ls_run_lmv": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f64, _f64, _f64, _f64, _f64, _i32, _f64,
                          C.POINTER(_f64), C.POINTER(_i32), C.POINTER(_i32), _vp],
    "more3d_ls_phase": [_vp, _i32, _vp, _vp],
    "more3d_ls_smooth_raw": [_vp, _i32, _vp, _vp, _i64, _i32, _vp, _vp, _vp],
    "more3d_ls_elementwise": [_i32, _vp, _f64, _i64, _vp, _vp],
    "more3d_normals_from_knn": [_vp, _i64, _i32, _vp, _vp, _vp, _vp],
    "more3d_orient_normals": [_vp, _i64, _i32, _vp],
    "more3d_compare_terms": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "more3d_band_select": [_vp, _i64, _i32, _f64, _f64, _vp, C.POINTER(_i64), _vp],
    "more3d_gather_rows": [_vp, _vp, _i64, _i32, _i32, _vp, _vp],
    "more3d_column_stats": [_vp, _i64, _vp, _vp],
    "more3d_select_kth": [_vp, _i64, C.POINTER(_i64), _i32, _vp, _vp],
    "more3d_cue_ops": [_i32, _vp, _i64, _vp, _f64, _vp, _vp],
    "more3d_saliency_blend_owned": [_vp, _vp, _vp, _vp, _f64, _i64, _i64, _vp, _vp, _vp],
    "more3d_strip_pack": [_vp, _i64, _f64, _f64, _vp, _vp, _i64, _vp, _vp],
    "more3d_histogram256_range": [_vp, _i32, _i64, _vp, _i32, _vp, _vp],
    "more3d_negate_odd": [_vp, _i64, _vp],
    "more3d_synth_terrain": [C.c_uint64, _i64, _i64, _i64, _f64, _f64, _i32, _f64, _vp, _vp],
}


class More3DError(RuntimeError):
    """Raised for every non-zero return code of the C-ABI; ``rc`` is that code (MORE3D_E_*)."""

    def __init__(self, msg, rc=None):
        super().__init__(msg)
        self.rc = rc


E_INVALID, E_CUDA, E_NOMEM, E_RANGE = -1, -2, -3, -4


def load(build_if_needed=True):
    """Load (building first when stale and nvcc is available) and return the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_needed and os.path.exists("/usr/local/cuda/bin/nvcc"):
        from . import build as _build
        try:
            if _build.stale():
                _build.build()
        except Exception as e:  # stale lib + failed rebuild must not be silently used
            raise More3DError("building libmore3d_b200.so failed: %s" % e)
    if not os.path.exists(LIB_PATH):
        raise More3DError("libmore3d_b200.so not found (run `python -m more3d_b200.build`); "
                          "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.more3d_last_error.restype = C.c_char_p
    lib.more3d_last_error.argtypes = []
    lib.more3d_version.restype = C.c_int
    lib.more3d_version.argtypes = []
    lib.more3d_launch_count.restype = C.c_int64
    lib.more3d_launch_count.argtypes = []
    lib.more3d_ls_scratch_doubles.restype = C.c_int64
    lib.more3d_ls_scratch_doubles.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue  # reported by tests/test_abi.py; entry points may land incrementally
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().more3d_last_error().decode("utf-8", "replace")
        raise More3DError("libmore3d_b200 error %d: %s" % (rc, msg), rc)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise More3DError("more3d_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def profile_enable(on=True):
    check(load().more3d_profile_enable(1 if on else 0))


def read_back(t):
    """Small CUDA tensor -> numpy array on the host, ordered behind the current stream's work, without the copy engine
    (more3d_read_back): a `.cpu()` of a few bytes waits for any large download that is in flight."""
    import numpy as np
    import torch
    t = t.contiguous()
    out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    with torch.cuda.device(t.device):
        check(load().more3d_read_back(ptr(t), out.ctypes.data_as(C.c_void_p), int(t.numel() * t.element_size()), stream_ptr()))
    return out


def set_dkdn_variant(variant):
    """0 = two-pass dk/dn (default), 1 = shared-memory staged walk, 2 = one-pass walk; applies to grids built afterwards."""
    check(load().more3d_set_dkdn_variant(int(variant)))


def set_dkdn_gate(fraction):
    """Listed fraction of the cloud above which the two-pass dk/dn falls back to the one-pass kernel (default 0.45)."""
    check(load().more3d_set_dkdn_gate(float(fraction)))


def profile_reset():
    check(load().more3d_profile_reset())


def profile_get(name):
    """(total ms, launches) of one kernel group since the last reset (synchronises its events)."""
    ms, cnt = C.c_double(0.0), C.c_int64(0)
    check(load().more3d_profile_get(name.encode(), C.byref(ms), C.byref(cnt)))
    return float(ms.value), int(cnt.value)


def profile_all():
    buf = C.create_string_buffer(4096)
    check(load().more3d_profile_names(buf, 4096))
    names = [n for n in buf.value.decode().split(",") if n]
    return {n: profile_get(n) for n in names}


def cloud_bounds(xyz):
    """(min x, y, z, max x, y, z) of an (N, 3) f64 device tensor as host floats -- the bounding-box reduction of
    more3d_strip_pack with both halo bands empty (one kernel, one small read)."""
    import torch
    n = int(xyz.shape[0])
    with torch.cuda.device(xyz.device):
        info = torch.empty(8, dtype=torch.float64, device=xyz.device)
        dummy = torch.empty(3, dtype=torch.float64, device=xyz.device)
        check(load().more3d_strip_pack(ptr(xyz), n, float("-inf"), float("inf"), ptr(dummy), ptr(dummy), 0, ptr(info),
                                       stream_ptr()))
    return info[:6].cpu().numpy()


def launch_count():
    return int(load().more3d_launch_count())
