"""Diagnostic: which host call stalls in the tile stream?  Wraps candidate functions and prints every call > 8 ms."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic, _lib, pipeline, grid, DM_saliency
n = 10_000_000
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pts = synthetic.terrain(n, seed=2)
host = torch.from_numpy(pts).pin_memory()
out_host = {r: torch.empty(n, dtype=torch.float32).pin_memory() for r in (0.5, 0.75, 1.0)}
P = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
lib = _lib.load()
pipeline.multiscale_saliency_tiles([host] * 3, out_hosts=[out_host] * 3, keep_results=False, **P)
torch.cuda.synchronize()
log = []
T0 = [0.0]
def wrap(obj, name, label=None):
    f = getattr(obj, name)
    def w(*a, **kw):
        t = time.perf_counter()
        r = f(*a, **kw)
        d = time.perf_counter() - t
        if d > 0.008:
            log.append(((t - T0[0]) * 1e3, d * 1e3, label or name))
        return r
    setattr(obj, name, w)
wrap(torch.Tensor, "pin_memory")
wrap(torch, "empty", "torch.empty")
wrap(torch.Tensor, "copy_")
wrap(torch.cuda.Event, "synchronize", "Event.synchronize")
wrap(torch.cuda.Stream, "synchronize", "Stream.synchronize")
wrap(DM_saliency.PointCloudDM, "check_pending_async")
wrap(DM_saliency.PointCloudDM, "drop_grids")
wrap(grid.UniformGrid, "normals_curvature_gridorder")
wrap(grid.UniformGrid, "dk_dn_gridorder")
wrap(grid.UniformGrid, "close")
# the C calls of the grid build
for nm in ("more3d_grid_create", "more3d_grid_create_segments", "more3d_grid_destroy", "more3d_strip_pack", "more3d_unpermute",
           "more3d_saliency_blend"):
    if hasattr(lib, nm):
        f = getattr(lib, nm)
        def mk(f, nm):
            def w(*a):
                t = time.perf_counter(); r = f(*a); d = time.perf_counter() - t
                if d > 0.008: log.append(((t - T0[0]) * 1e3, d * 1e3, nm))
                return r
            return w
        setattr(lib, nm, mk(f, nm))
T0[0] = time.perf_counter()
pipeline.multiscale_saliency_tiles([host] * k, out_hosts=[out_host] * k, keep_results=False, **P)
print("total ms", (time.perf_counter() - T0[0]) * 1e3)
for t, d, nm in log:
    print("%8.1f  %7.1f ms  %s" % (t, d, nm))
