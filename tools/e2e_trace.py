"""Per-tile timeline of the tile-streaming e2e path (diagnostic): host time at which each tile was enqueued and
CUDA-event time at which its kernels / its download finished."""
import sys, os, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic, _lib, pipeline
n = 10_000_000
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pts = synthetic.terrain(n, seed=2)
host = torch.from_numpy(pts).pin_memory()
out_host = {r: torch.empty(n, dtype=torch.float32).pin_memory() for r in (0.5, 0.75, 1.0)}
P = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
_lib.load()
pipeline.multiscale_saliency_tiles([host] * 3, out_hosts=[out_host] * 3, keep_results=False, **P)
torch.cuda.synchronize()
# instrument: wrap multiscale_saliency to record host time + event per tile
orig = pipeline.multiscale_saliency
marks = []
def wrapped(*a, **kw):
    t0 = time.perf_counter()
    r = orig(*a, **kw)
    ev = torch.cuda.Event(enable_timing=True); ev.record()
    marks.append((t0, time.perf_counter(), ev))
    return r
pipeline.multiscale_saliency = wrapped
start = torch.cuda.Event(enable_timing=True); start.record()
T0 = time.perf_counter()
pipeline.multiscale_saliency_tiles([host] * k, out_hosts=[out_host] * k, keep_results=False, **P)
T1 = time.perf_counter()
torch.cuda.synchronize()
print("total wall ms", (T1 - T0) * 1e3, "mem GB", torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
prev = 0.0
for i, (a, b, ev) in enumerate(marks):
    g = start.elapsed_time(ev)
    print(i, "host enqueue start %.2f dur %.2f | gpu done %.2f (+%.2f)" % ((a - T0) * 1e3, (b - a) * 1e3, g, g - prev))
    prev = g

# ---- second pass: where does the host time of a slow tile go?  (cProfile over one 20-tile stream)
if len(sys.argv) > 2:
    import cProfile, pstats, io
    pipeline.multiscale_saliency = orig
    pr = cProfile.Profile()
    pr.enable()
    pipeline.multiscale_saliency_tiles([host] * k, out_hosts=[out_host] * k, keep_results=False, **P)
    pr.disable()
    st = io.StringIO()
    pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(18)
    print(st.getvalue()[:5000])
