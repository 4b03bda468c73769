import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from more3d_b200 import synthetic, _lib
from more3d_b200.distributed import strip_pack
pts = torch.from_numpy(synthetic.terrain(10_000_000, seed=2)).cuda()
for _ in range(3): strip_pack(pts, -248.0, 248.0, 200000)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): l, r, info = strip_pack(pts, -248.0, 248.0, 200000)
b.record(); torch.cuda.synchronize()
print("strip_pack ms", a.elapsed_time(b) / 20, info.cpu().numpy())
