"""Timings of the index-construction kernels at the sizes VERDICT r01 item 2 names (CUDA events, best of 3):
grid build at 10 M (host-order terrain) and 62.5 M points, voxel grid at 50 M, LOD tree at 50 M."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic, _lib  # noqa: E402
from more3d_b200.grid import UniformGrid, cell_for_radius  # noqa: E402
from more3d_b200.Downsample import voxel_downsample  # noqa: E402
from more3d_b200.lodtree import LodTree  # noqa: E402
from bench_configs import best_ms  # noqa: E402

out = {}
_lib.load()
d = torch.from_numpy(synthetic.terrain(10_000_000, seed=2)).cuda()
_lib.profile_enable(True)
t, g = best_ms(lambda: UniformGrid(d, cell_for_radius(0.5)))
out["grid_build_10M_random_order_ms"] = t
g.close()
del d, g
d = synthetic.hash_terrain_device(5, 0, 62_500_000, 500_000_000)
t, g = best_ms(lambda: UniformGrid(d, cell_for_radius(0.5)))
out["grid_build_62.5M_slab_order_ms"] = t
g.close()
del d, g
torch.cuda.empty_cache()
d = synthetic.hash_terrain_device(3, 0, 50_000_000, 50_000_000)
for lod in (1, 2, 5, 10):
    _lib.profile_reset()
    t, v = best_ms(lambda: voxel_downsample(d, float(lod), return_device=True), reps=3)
    prof = _lib.profile_all()
    out["voxel_50M_lod%d_ms" % lod] = {"total": t, "voxels": int(v.shape[0]),
                                        "parts_ms_per_call": {k: a / max(c, 1) for k, (a, c) in prof.items() if k.startswith("voxel")}}
    del v
t, tr = best_ms(lambda: LodTree(d, 40), reps=2)
out["tree_build_50M_ms"] = t
print(json.dumps(out))
