"""Two saliency passes per radius on 10 M points (ncu target: `-k regex:dkdn -c 6` captures both dk/dn launches of
r = 0.5, 0.75, 1.0; MORE3D_DKDN_VARIANT selects the kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic  # noqa: E402
from more3d_b200.pipeline import multiscale_saliency  # noqa: E402

PARAMS = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
d = synthetic.hash_terrain_device(2, 0, 10_000_000, 10_000_000)
for r in (0.5, 0.75, 1.0):
    for _ in range(2):
        multiscale_saliency(d, (r,), **PARAMS)
print("done")
