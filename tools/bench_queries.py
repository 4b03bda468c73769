"""Timing of EXTERNAL query sets (queries != the cloud): radius count + fill and kNN, shuffled queries.
Run on the GPU box: python tools/bench_queries.py [n_points] [n_queries]"""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from more3d_b200 import synthetic
from more3d_b200.grid import UniformGrid, cell_for_radius

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
pts = torch.from_numpy(synthetic.terrain(n, seed=1)).cuda()
g = UniformGrid(pts, cell_for_radius(0.5))
perm = torch.randperm(n, device="cuda")[:nq]
q = (pts[perm] + 0.05 * torch.randn(nq, 3, device="cuda", dtype=torch.float64)).contiguous()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print("points", n, "queries", nq, "(shuffled)")
print("radius_count r=0.5   %.2f ms" % timed(lambda: g.radius_count(q, 0.5)))
print("radius_csr   r=0.5   %.2f ms" % timed(lambda: g.radius_csr(q, 0.5)))
print("knn k=8              %.2f ms" % timed(lambda: g.knn(q, 8)))
print("knn k=16             %.2f ms" % timed(lambda: g.knn(q, 16)))
