"""One pass of every index-construction path at its reference size, for an ncu launch list
(`ncu --metrics gpu__time_duration.sum ... python tools/prof_build.py`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic, _lib  # noqa: E402
from more3d_b200.grid import UniformGrid, cell_for_radius  # noqa: E402
from more3d_b200.Downsample import voxel_downsample  # noqa: E402
from more3d_b200.lodtree import LodTree  # noqa: E402

_lib.load()
n = int(os.environ.get("PROF_N", "50000000"))
d = torch.from_numpy(synthetic.terrain(10_000_000, seed=2)).cuda()
for _ in range(2):
    UniformGrid(d, cell_for_radius(0.5)).close()
del d
d = synthetic.hash_terrain_device(3, 0, n, n)
for _ in range(2):
    voxel_downsample(d, 1.0, return_device=True)
LodTree(d, 40)
torch.cuda.synchronize()
print("done")
