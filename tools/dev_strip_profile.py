"""Where does a strip step spend its time beyond the neighbourhood kernels?  (development aid, one GPU, one-rank NCCL)"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more3d_b200 import distributed as D

RADII = (0.5, 0.75, 1.0)
PARAMS = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29633", rank=0, world_size=1, device_id=dev)
job = D.StripSaliency(int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000, 0, 1, dev, RADII, PARAMS, 0.5)
for _ in range(3):
    job.step_resident()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


torch.cuda.synchronize()
own = job.resident
t0 = time.time(); e0 = ev()
left = D.gather_rows(own, D.band_select(own, 0, -np.inf, job.x_lo + job.halo))
right = D.gather_rows(own, D.band_select(own, 0, job.x_hi - job.halo, np.inf))
e1 = ev()
fl, fr = D.exchange_halos(left, right, 0, 1)
ghosts = torch.cat([fl, fr], dim=0)
e2 = ev()
out, pc = D.strip_saliency(own, ghosts, RADII, PARAMS, 0.5, reduce_minmax=D.allreduce_minmax)
e3 = ev()
jobs = D.DeviceOtsu([out[r] for r in RADII])
res = jobs.results()
e4 = ev()
torch.cuda.synchronize(); t1 = time.time()
print("select %.3f | exchange %.3f | strip_saliency %.3f | otsu %.3f | total gpu %.3f | wall %.3f ms"
      % (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4), e0.elapsed_time(e4),
         (t1 - t0) * 1e3))
torch.cuda.synchronize()
a = ev()
for _ in range(5):
    job.step_resident()
b = ev(); torch.cuda.synchronize()
print("step_resident %.3f ms" % (a.elapsed_time(b) / 5))
from more3d_b200.pipeline import multiscale_saliency
a = ev()
for _ in range(5):
    multiscale_saliency(own, RADII, curvature_weight=0.5, **PARAMS)
b = ev(); torch.cuda.synchronize()
print("multiscale_saliency %.3f ms" % (a.elapsed_time(b) / 5))
dist.destroy_process_group()
