"""BASELINE.json configs[4]: a 500 M-point tile, saliency pipeline at r = 0.5, strong scaling.

    torchrun --nproc-per-node 8 bench.py --gpus 8 --points 62500000 --radii 0.5      # 8 GPUs, 62.5 M points each
    python tools/bench_config5.py 8 62500000                                    # the SAME tile on one GPU

The single-GPU run concatenates the eight strips exactly as the ranks generate them, so both runs process
identical points; it prints the time of the whole-tile pipeline and the neighbour-count checksum.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic  # noqa: E402
from more3d_b200.pipeline import multiscale_saliency  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 62_500_000
W = float(np.sqrt(n / synthetic.DENSITY))
t0 = time.time()
dev = torch.empty((G * n, 3), dtype=torch.float64, device="cuda")
for k in range(G):
    p = synthetic.terrain(n, seed=2 + k, extent=(W, W), origin=(k * W, 0.0), centre=(0.5 * G * W, 0.5 * W))
    dev[k * n:(k + 1) * n] = torch.from_numpy(p).cuda()
    del p
gen_s = time.time() - t0
params = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
res = multiscale_saliency(dev, (0.5,), curvature_weight=0.5, keep=True, **params)
sum_m = int(res[0.5]["_pcount"].long().sum())
del res
torch.cuda.synchronize()
times = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = multiscale_saliency(dev, (0.5,), curvature_weight=0.5, **params)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
    del out
print(json.dumps({"config": 5, "gpus": 1, "points": G * n, "radius": 0.5, "ms_per_step": min(times), "all_ms": times,
                  "points_per_s": G * n / (min(times) * 1e-3), "sum_pcount": sum_m, "host_generation_s": gen_s,
                  "hbm_allocated_GB": torch.cuda.max_memory_allocated() / 1e9}))
