"""Seeded synthetic open-terrain clouds.

All inputs are generated on the host in fp64 with ``numpy.random.Generator(PCG64(seed))``:
xy ~ U(0, L)^2 with L = sqrt(N / density); z = 7-octave sin*cos relief + N(0, noise^2) + one
Gaussian bump/pit per 50 m x 50 m block; finally x, y are centred and z is shifted by -1000 m so
that the coordinate origin ("scanner") sits above the tile -- ``Curvature_Kernel`` orients normals
toward the origin (reference DM_saliency.py:131-135).
"""
import numpy as np

DENSITY = 40.0  # points / m^2


def terrain(n, seed, density=DENSITY, noise=0.02, boulders=0, extent=None, origin=(0.0, 0.0), centre=None):
    """Return an (n, 3) float64 C-order terrain cloud.

    ``boulders`` > 0 adds that many steep 0.5 m hemispherical bumps per 50 m block (the "stress
    tile" of SURVEY.md §8d that exercises the dn branch of the saliency processor).
    ``extent`` = (Lx, Ly) overrides the square footprint; ``origin`` = (x0, y0) places the footprint
    in a larger tile (the relief is a function of the global position, so neighbouring strips join
    continuously) and ``centre`` = the global point moved to x = y = 0 (default: footprint centre).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if extent is None:
        L = float(np.sqrt(n / density))
        Lx = Ly = L
    else:
        Lx, Ly = map(float, extent)
    x = rng.random(n) * Lx
    y = rng.random(n) * Ly
    z = np.zeros(n)
    for o in range(7):
        f = float(2 ** o)
        z += 2.0 * 2.0 ** (-0.8 * o) * np.sin(f * (x + origin[0]) / 40.0 + o) * np.cos(f * (y + origin[1]) / 37.0 + 2 * o)
    z += rng.normal(0.0, noise, n)

    # embedded entities: one Gaussian bump / pit per 50 m block, centre kept >= 10 m from the
    # block border so a point is only influenced by the entity of its own block.
    B = 50.0
    nbx = max(1, int(np.ceil(Lx / B)))
    nby = max(1, int(np.ceil(Ly / B)))
    erng = np.random.Generator(np.random.PCG64(seed + 7919))
    cx = (np.arange(nbx)[:, None] * B + 10.0 + 30.0 * erng.random((nbx, nby)))
    cy = (np.arange(nby)[None, :] * B + 10.0 + 30.0 * erng.random((nbx, nby)))
    amp = 0.4 * np.where(erng.random((nbx, nby)) < 0.5, -1.0, 1.0)
    sig = 1.0 + erng.random((nbx, nby))
    bx = np.minimum((x / B).astype(np.int64), nbx - 1)
    by = np.minimum((y / B).astype(np.int64), nby - 1)
    d2 = (x - cx[bx, by]) ** 2 + (y - cy[bx, by]) ** 2
    z += amp[bx, by] * np.exp(-0.5 * d2 / sig[bx, by] ** 2)

    if boulders:
        R = 0.5
        for _ in range(int(boulders)):
            ox = (np.arange(nbx)[:, None] * B + 2.0 + 46.0 * erng.random((nbx, nby)))
            oy = (np.arange(nby)[None, :] * B + 2.0 + 46.0 * erng.random((nbx, nby)))
            d2 = (x - ox[bx, by]) ** 2 + (y - oy[bx, by]) ** 2
            z += np.sqrt(np.maximum(R * R - d2, 0.0))

    pts = np.empty((n, 3), dtype=np.float64)
    if centre is None:
        centre = (origin[0] + 0.5 * Lx, origin[1] + 0.5 * Ly)
    pts[:, 0] = (x + origin[0]) - centre[0]
    pts[:, 1] = (y + origin[1]) - centre[1]
    pts[:, 2] = z - 1000.0
    return pts


def lattice(nx, ny, nz, spacing=1.0):
    """Integer lattice: exact radius ties (SURVEY.md §8c known-answer test)."""
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1)
    return np.ascontiguousarray(g.reshape(-1, 3).astype(np.float64) * spacing)


def plane(n, seed, normal=(0.0, 0.0, 1.0), extent=10.0, z0=-1000.0):
    """Random points on an exact plane through (0,0,z0) (K = 0, dn = dk = 0 known answers)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nrm = np.asarray(normal, dtype=np.float64)
    nrm = nrm / np.linalg.norm(nrm)
    a = np.cross(nrm, [1.0, 0.0, 0.0])
    if np.linalg.norm(a) < 1e-6:
        a = np.cross(nrm, [0.0, 1.0, 0.0])
    a /= np.linalg.norm(a)
    b = np.cross(nrm, a)
    uv = (rng.random((n, 2)) - 0.5) * extent
    pts = uv[:, :1] * a[None, :] + uv[:, 1:] * b[None, :]
    pts[:, 2] += z0
    return np.ascontiguousarray(pts)


def sphere_cap(n, seed, radius=50.0, cap=0.2):
    """Points on the top cap of a sphere centred at the origin-shifted location (0, 0, -1000 - R)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(n)
    cos_t = 1.0 - u * (1.0 - np.cos(cap))
    sin_t = np.sqrt(1.0 - cos_t ** 2)
    ph = rng.random(n) * 2 * np.pi
    pts = np.stack([radius * sin_t * np.cos(ph), radius * sin_t * np.sin(ph),
                    radius * cos_t - radius - 1000.0], -1)
    return np.ascontiguousarray(pts)


# ------------------------------------------------------------------------------------------------
# Counter-based terrain: point i of a tile is a pure function of (seed, i) -- no sequential RNG stream -- so
# any rank can generate any id range of a 50 M / 500 M-point tile, on the DEVICE (`more3d_synth_terrain`,
# csrc/synth.cu; milliseconds instead of minutes of host numpy), and every partition of the ids over 1/2/4/8
# GPUs is the SAME global cloud (bench.py's strong-scaling leg compares checksums across partitions).
# The tile's x range is cut into `slabs` equal slabs; ids are slab-major (slab = i * slabs // total), uniform
# inside the slab, so an id range [k * total / G, (k + 1) * total / G) is exactly the k-th of G x-strips when
# G divides `slabs`.  Same relief / noise / entity model as `terrain` above.
# `hash_terrain_host` is the numpy twin of the kernel (integer hashing identical; sin/cos/exp/log differ by
# a few ulp between libm and CUDA, so z agrees to ~1e-12 m, x and y bit for bit).
# ------------------------------------------------------------------------------------------------
HASH_SLABS = 4000
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _u01(seed, ids, stream):
    """53-bit uniform in [0, 1) from (seed, id, stream)."""
    with np.errstate(over="ignore"):
        k = _splitmix64(_splitmix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))) ^ ids.astype(np.uint64))
    return (k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hash_tile_extent(total, density=DENSITY, aspect=1.0):
    """(Lx, Ly) of a tile of `total` points at `density`, Lx = aspect * Ly."""
    area = total / density
    ly = float(np.sqrt(area / aspect))
    return aspect * ly, ly


def hash_terrain_host(seed, id_start, count, total, extent=None, noise=0.02, slabs=HASH_SLABS):
    """numpy twin of more3d_synth_terrain: points [id_start, id_start + count) of the `total`-point tile."""
    lx, ly = hash_tile_extent(total) if extent is None else map(float, extent)
    ids = np.arange(id_start, id_start + count, dtype=np.uint64)
    # slab = floor(id * slabs / total) without 128-bit products: total / slabs points per slab when divisible
    slab = ((ids.astype(np.float64) + 0.5) * (slabs / float(total))).astype(np.int64)
    x = (slab + _u01(seed, ids, 1)) * (lx / slabs)
    y = _u01(seed, ids, 2) * ly
    z = np.zeros(count)
    for o in range(7):
        f = float(2 ** o)
        z += 2.0 * 2.0 ** (-0.8 * o) * np.sin(f * x / 40.0 + o) * np.cos(f * y / 37.0 + 2 * o)
    u1, u2 = _u01(seed, ids, 3), _u01(seed, ids, 4)
    z += noise * np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
    B = 50.0
    bx, by = np.floor(x / B), np.floor(y / B)
    bid = (bx.astype(np.int64) * 1000003 + by.astype(np.int64)).astype(np.uint64)
    cx = bx * B + 10.0 + 30.0 * _u01(seed, bid, 5)
    cy = by * B + 10.0 + 30.0 * _u01(seed, bid, 6)
    amp = np.where(_u01(seed, bid, 7) < 0.5, -0.4, 0.4)
    sig = 1.0 + _u01(seed, bid, 8)
    z += amp * np.exp(-0.5 * ((x - cx) ** 2 + (y - cy) ** 2) / (sig * sig))
    pts = np.empty((count, 3))
    pts[:, 0] = x - 0.5 * lx
    pts[:, 1] = y - 0.5 * ly
    pts[:, 2] = z - 1000.0
    return pts


def hash_terrain_device(seed, id_start, count, total, extent=None, noise=0.02, slabs=HASH_SLABS, device=None, out=None):
    """Points [id_start, id_start + count) of the tile, generated on the GPU: (count, 3) float64 CUDA tensor."""
    import ctypes as C
    import torch
    from . import _lib
    _lib.require_cuda()
    lx, ly = hash_tile_extent(total) if extent is None else map(float, extent)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if out is None:
        out = torch.empty((int(count), 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().more3d_synth_terrain(C.c_uint64(int(seed)), int(id_start), int(count), int(total),
                                                    float(lx), float(ly), int(slabs), float(noise), _lib.ptr(out),
                                                    _lib.stream_ptr()))
    return out
