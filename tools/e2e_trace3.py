"""Diagnostic: CUDA-event timeline of the tile stream -- per tile: upload [start, end] on the upload stream, kernels
[start, end] on the compute stream, first download start / last download end on the download stream."""
import sys, os, time, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from more3d_b200 import synthetic, _lib, pipeline
n = 10_000_000
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
smi = len(sys.argv) > 2 and sys.argv[2] == "smi"
pts = synthetic.terrain(n, seed=2)
host = torch.from_numpy(pts).pin_memory()
out_host = {r: torch.empty(n, dtype=torch.float32).pin_memory() for r in (0.5, 0.75, 1.0)}
P = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
_lib.load()
pipeline.multiscale_saliency_tiles([host] * 3, out_hosts=[out_host] * 3, keep_results=False, **P)
torch.cuda.synchronize()
proc = None
if smi:
    proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                             "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    time.sleep(0.5)
E = lambda: torch.cuda.Event(enable_timing=True)
dev = torch.device("cuda", 0)
cur = torch.cuda.current_stream(dev)
up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
bufs = pipeline._staging_buffers(dev, n)
freed = [None, None]
rec = []
def upload(i):
    slot = i % 2
    with torch.cuda.stream(up):
        if freed[slot] is not None:
            up.wait_event(freed[slot])
        else:
            up.wait_stream(cur)
        a = E(); a.record(up)
        bufs[slot].copy_(host, non_blocking=True)
        b = E(); b.record(up)
    return bufs[slot], a, b
t0 = E(); t0.record(cur)
W0 = time.perf_counter()
nxt = upload(0)
checks = []
for i in range(k):
    d, ua, ub = nxt
    if i + 1 < k:
        nxt = upload(i + 1)
    cur.wait_event(ub)
    ka = E(); ka.record(cur)
    da = E()
    with torch.cuda.stream(down):
        da.record(down)
    pend = []
    h0 = time.perf_counter()
    pipeline.multiscale_saliency(d, (0.5, 0.75, 1.0), out_host=out_host, copy_stream=down, deferred=pend, **P)
    h1 = time.perf_counter()
    kb = E(); kb.record(cur)
    db = E()
    with torch.cuda.stream(down):
        db.record(down)
    freed[i % 2] = kb
    checks.append(pend[0])
    if i > 0:
        checks[i - 1]()
    rec.append((ua, ub, ka, kb, db, (h0 - W0) * 1e3, (h1 - h0) * 1e3))
cur.wait_stream(down)
cur.synchronize()
print("total wall ms %.1f  per tile %.2f" % ((time.perf_counter() - W0) * 1e3, (time.perf_counter() - W0) * 1e3 / k))
for i, (ua, ub, ka, kb, db, hs, hd) in enumerate(rec):
    f = lambda e: t0.elapsed_time(e)
    print("%2d upload %.1f-%.1f (%.1f) | kernels %.1f-%.1f (%.1f) | last D2H done %.1f | host enqueue at %.1f took %.1f"
          % (i, f(ua), f(ub), f(ub) - f(ua), f(ka), f(kb), f(kb) - f(ka), f(db), hs, hd))
if proc:
    proc.terminate()
