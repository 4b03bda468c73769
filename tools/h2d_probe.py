"""Host <-> device copy rates of this box from pinned memory (what bounds bench.py's e2e leg): H2D of a 240 MB tile,
D2H of 3 x 40 MB columns, alone and together.  Prints one JSON line."""
import json
import time

import torch

n = 10_000_000
dev = torch.device("cuda", 0)
host = torch.empty((n, 3), dtype=torch.float64).pin_memory()
host.fill_(1.0)
out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
d = torch.empty((n, 3), dtype=torch.float64, device=dev)
cols = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(3)]
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def h2d():
    with torch.cuda.stream(up):
        d.copy_(host, non_blocking=True)


def d2h():
    with torch.cuda.stream(down):
        for c, o in zip(cols, out):
            o.copy_(c, non_blocking=True)


def both():
    h2d()
    d2h()


res = {}
t = timed(h2d); res["h2d_GBs"] = 0.24 / t
t = timed(d2h); res["d2h_GBs"] = 0.12 / t
t = timed(both); res["both_ms"] = t * 1e3
pg = torch.empty((n, 3), dtype=torch.float64)
pg.fill_(1.0)
t = timed(lambda: d.copy_(pg)); res["h2d_pageable_GBs"] = 0.24 / t
print(json.dumps(res))
