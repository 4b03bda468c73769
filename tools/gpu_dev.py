"""Per-stage CUDA-event timings of the saliency pipeline (development aid; run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic  # noqa: E402
from more3d_b200.grid import UniformGrid, CELL_MARGIN  # noqa: E402
from more3d_b200.DM_saliency import _chi2_table  # noqa: E402


def timed(fn, reps=3):
    out = None
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


DIV = int(os.environ.get("MORE3D_CELLS_PER_RADIUS", "2"))


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
    radii = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.5, 0.75, 1.0]
    t0 = time.time()
    pts = synthetic.terrain(n, seed=2)
    print("gen %.1fs" % (time.time() - t0), flush=True)
    d = torch.from_numpy(pts).cuda()
    table = _chi2_table(0.05, 4096, d.device)
    for r in radii:
        cell = float(os.environ["MORE3D_FIXED_CELL"]) if "MORE3D_FIXED_CELL" in os.environ else r * CELL_MARGIN / DIV
        tb, g = timed(lambda: UniformGrid(d, cell))
        tc, cnt = timed(lambda: g.radius_count(None, r))
        tm, (nrm, ratio, K, pc) = timed(lambda: g.normals_curvature(r, 0.0392, 4))
        td, (dk, dn, mm) = timed(lambda: g.dk_dn(r, nrm, K, r / np.sqrt(2), 0.01, 0.01, table, 4))
        m = float(pc.double().mean())
        print("N=%d r=%.2f grid=%dx%d  build %.2f ms | count %.2f ms | moments %.2f ms | dkdn %.2f ms | mean m %.1f"
              % (n, r, g.nx, g.ny, tb, tc, tm, td, m), flush=True)
        a_mom = n * 48 + 16 * m * n
        a_dk = n * 40 + 32 * m * n
        print("   algorithmic GB/s: moments %.0f  dkdn %.0f" % (a_mom / tm / 1e6, a_dk / td / 1e6), flush=True)
        g.close()


if __name__ == "__main__":
    main()
