"""A/B of the dk/dn kernel variants on configs[1] (10 M points, 3 radii): per-radius CUDA-event times of the dkdn and
moments groups, and bit-equality of the outputs against variant 0.  Run once per variant:
    MORE3D_DKDN_VARIANT=1 MORE3D_SMEM_CAP=640 python tools/ab_dkdn.py out.npz [ref.npz]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic, _lib  # noqa: E402
from more3d_b200.pipeline import multiscale_saliency  # noqa: E402

PARAMS = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
n = int(os.environ.get("AB_N", "10000000"))
d = synthetic.hash_terrain_device(2, 0, n, n)
_lib.load()
out = {"variant": os.environ.get("MORE3D_DKDN_VARIANT", "0"), "cap": os.environ.get("MORE3D_SMEM_CAP", ""), "n": n}
for r in (0.5, 0.75, 1.0):
    for _ in range(2):
        multiscale_saliency(d, (r,), **PARAMS)
    _lib.profile_reset()
    _lib.profile_enable(True)
    reps = 3
    for _ in range(reps):
        multiscale_saliency(d, (r,), **PARAMS)
    prof = _lib.profile_all()
    _lib.profile_enable(False)
    out[str(r)] = {k: round(v[0] / reps, 4) for k, v in prof.items() if k in ("dkdn", "dkdn_screen", "dkdn_refine", "moments", "grid_build")}
res = multiscale_saliency(d, (0.5, 0.75, 1.0), keep=True, **PARAMS)
arrs = {}
for r in (0.5, 0.75, 1.0):
    for k in ("_dk", "_dn", "_saliency", "_nonparam_K"):
        arrs["%s%s" % (r, k)] = res[r][k].cpu().numpy()
if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
    ref = np.load(sys.argv[2])
    out["bit_identical_to_ref"] = bool(all(np.array_equal(arrs[k], ref[k], equal_nan=True) for k in arrs))
    # variants that add in another order (the two-pass kernel's refinement): same zeros / NaNs, values to rounding
    cmp = {}
    for k in arrs:
        a, b = arrs[k].astype(np.float64), ref[k].astype(np.float64)
        same_class = bool(np.array_equal(a == 0, b == 0) and np.array_equal(np.isnan(a), np.isnan(b)))
        nz = (b != 0) & ~np.isnan(b) & ~np.isnan(a)
        rel = float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz]))) if nz.any() else 0.0
        cmp[k] = {"same_zero_nan_pattern": same_class, "max_rel_diff": rel, "n_diff": int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b)))))}
    out["compare"] = cmp
if len(sys.argv) > 1:
    np.savez(sys.argv[1], **arrs)
print(json.dumps(out))
