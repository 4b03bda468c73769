"""Level-set iteration timing on a synthetic terrain (development aid; run under gpurun, optionally under ncu)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more3d_b200 import synthetic, levelsets_func as ls  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pts = synthetic.terrain(n, seed=4)
pts -= pts.mean(axis=0)
rng = np.random.Generator(np.random.PCG64(1))
zeta = (rng.random(n) ** 6)[:, None]
h, k = 0.5, 7
nb, nrm, tang = ls.build_neighborhoods(pts, h, k, return_device=True)
solver = ls.Solver(h, zeta, pts, nb, tang, "chan_vese")
solver.initialize_from_voxels(2.0)
run = dict(stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
solver.run(5, **run)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
solver.run(iters, **run)
b.record()
torch.cuda.synchronize()
print("n=%d  %.3f ms/iteration  active=%d" % (n, a.elapsed_time(b) / iters, int(solver.active.sum())))
