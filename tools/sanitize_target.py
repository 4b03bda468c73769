"""Small end-to-end exercise of every kernel family (run under compute-sanitizer --tool memcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more3d_b200 import synthetic, levelsets_func as ls  # noqa: E402
from more3d_b200.BallTreePointSet import BallTreePointSet  # noqa: E402
from more3d_b200.Downsample import saliency_downsample, voxel_downsample  # noqa: E402
from more3d_b200.downsample_compare import compare  # noqa: E402
from more3d_b200.pipeline import multiscale_saliency  # noqa: E402

pts = synthetic.terrain(20011, seed=1)
rng = np.random.Generator(np.random.PCG64(1))
tall = np.vstack([pts, rng.random((9000, 3)) * [2.0, 2.0, 40.0] + [0, 0, -1020.0]])
for cloud in (pts, tall):
    res = multiscale_saliency(cloud, (0.5, 1.0), keep=True)
    bt = BallTreePointSet(cloud)
    bt.queryRadius(cloud[:777] + 0.01, 0.7)
    bt.query(cloud[:333] + 0.02, 9)
    sal = res[0.5]["_saliency"].cpu().numpy()
    sal = np.nan_to_num(sal)
    saliency_downsample(cloud, sal, 2.0, btP=bt)
    down = voxel_downsample(cloud, 0.8)
    compare(cloud, down, rnn=0.9)
c = pts - pts.mean(0)
nb, nrm, tang = ls.build_neighborhoods(c, 0.5, 7)
s = ls.Solver(0.5, np.abs(np.sin(c[:, :1])), c, nb, tang)
s.initialize_from_voxels(2.0)
s.run(6, stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
s.initialize_from_mask(c[:, 0] < 0)
s.run(4, stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
s.extract(4)
m = ls.Solver(0.5, np.abs(np.sin(c[:, :1])), c, nb, tang, "lmv")
m.initialize_from_voxels(2.0)
m.run(4, stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
import torch  # noqa: E402
from more3d_b200.pipeline import multiscale_saliency_tiles  # noqa: E402
tiles = [torch.from_numpy(pts[:15000]).pin_memory(), torch.from_numpy(pts[5000:]).pin_memory()]
outs = [{r: torch.empty(len(t), dtype=torch.float32).pin_memory() for r in (0.5, 0.75)} for t in tiles]
multiscale_saliency_tiles(tiles, (0.5, 0.75), out_hosts=outs)
print("sanitize target ok", float(np.abs(s.phi).max()), float(np.abs(m.phi).max()))
