"""Multi-GPU level set on a synthetic tile: one x-strip of n points per rank (run under torchrun).

    torchrun --nproc-per-node 2 tools/run_levelset_strips.py 10000000 50
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more3d_b200 import synthetic  # noqa: E402
from more3d_b200.distributed_levelset import StripLevelSet, DistComm  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = float(np.sqrt(n / synthetic.DENSITY))
pts = synthetic.terrain(n, seed=4 + rank, extent=(W, W), origin=(rank * W, 0.0), centre=(0.5 * world * W, 0.5 * W))
x_lo = rank * W - 0.5 * world * W
zeta = np.abs(np.sin(pts[:, :1] / 3.0) * np.cos(pts[:, 1:2] / 2.0))
h, k = 0.5, 7
s = StripLevelSet(DistComm(), pts, zeta, h, k, x_lo if rank > 0 else -np.inf, x_lo + W if rank < world - 1 else np.inf)
s.initialize_from_voxels(2.0)
s.smooth_phi()
run = dict(stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)
s.run(5, **run)
dist.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
s.run(iters, **run)
b.record()
dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
act = torch.tensor([float(s.active[:s.n].sum())], device="cuda", dtype=torch.float64)
dist.all_reduce(act)
if rank == 0:
    print(json.dumps({"gpus": world, "points_per_gpu": n, "iterations": iters, "ms_per_iteration": float(ms) / iters,
                      "point_iterations_per_s": n * world * iters / (float(ms) * 1e-3), "ghosts_rank0": s.nloc - s.n,
                      "active_total": float(act), "rmsc_last": s.rmsc_history[-1], "singular": s.singular}))
dist.destroy_process_group()
