"""The full DM_saliency pipeline at several neighbourhood radii (BASELINE.json configs[1]).

For every radius: DM_nonparam_curvature (normals at the same radius, fused) -> DM_dk_dn ->
DM_saliency, exactly the call sequence a user of the reference runs (README "Saliency"), on one
in-memory cloud.
"""
import numpy as np
import torch

from .DM_saliency import PointCloudDM, DM_nonparam_curvature, DM_dk_dn, DM_saliency

DEFAULT_RADII = (0.5, 0.75, 1.0)
DEFAULTS = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01,
                curvature_weight=0.5, min_points=4)


def multiscale_saliency(xyz, radii=DEFAULT_RADII, roughness=0.02, alpha=0.05, sigma_normal=0.01,
                        sigma_curvature=0.01, curvature_weight=0.5, min_points=4, keep=False,
                        out_host=None, copy_stream=None, chi2_table_len=1024, deferred=None):
    """Run the saliency pipeline at each radius (rho = r / sqrt(2)).

    xyz: (N,3) float64, numpy (host; copied to the device here) or a CUDA tensor.
    Returns {radius: saliency float32 CUDA tensor}; with ``keep`` also the intermediate columns.
    ``out_host``: optional {radius: pinned float32 host tensor} to receive the result (D2H copy).
    ``copy_stream``: the stream those D2H copies run on; when given, the caller waits for it (the compute stream
    does not, so the next tile's kernels start while the last column of this one is still travelling).
    The chi2-table overflow words of the dk/dn passes are read ONCE, after the last radius was queued (no host
    synchronisation between the radii); on overflow the pipeline is repeated with a table of the reported length.
    ``deferred``: a list -- the check is then left to the caller: a callable returning 0 or the needed table length
    is appended instead (tile streaming checks tile k while tile k + 1 computes).
    """
    dm = PointCloudDM(xyz)
    dm.defer_checks = True
    dm.chi2_table_len = int(chi2_table_len)
    out = {}
    own_copy_stream = copy_stream is None
    if out_host is not None and copy_stream is None:
        copy_stream = torch.cuda.Stream(device=dm.device)
    for r in sorted(radii):      # ascending: the finest grid is built first and shared by the wider radii
        rho = r / np.sqrt(2.0)
        dm.normals = None          # normals are re-estimated at every scale
        DM_nonparam_curvature(dm, "sphere", roughness=roughness, alpha=alpha, min_points=min_points, radius=r)
        DM_dk_dn(dm, "sphere", rho, rho, sigma_normal=sigma_normal, sigma_curvature=sigma_curvature,
                 alpha=alpha, min_points=min_points, radius=r)
        DM_saliency(dm, curvature_weight=curvature_weight)
        sal = dm.attrs["_saliency"]
        if out_host is not None:
            # device -> host on a side stream so the copy overlaps the next radius' kernels
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dm.device))
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev)
                out_host[r].copy_(sal, non_blocking=True)
            hold_until_copied(sal, copy_stream)
        if keep:
            out[r] = {k: v for k, v in dm.attrs.items()}
            out[r]["normals"] = dm.normals
        else:
            out[r] = sal
    if out_host is not None and own_copy_stream:
        torch.cuda.current_stream(dm.device).wait_stream(copy_stream)
    dm.drop_grids()
    if deferred is not None:
        side = copy_stream if copy_stream is not None else torch.cuda.Stream(device=dm.device)
        deferred.append(dm.check_pending_async(side))
        return out
    need = dm.check_pending()
    if need:
        return multiscale_saliency(xyz, radii, roughness, alpha, sigma_normal, sigma_curvature, curvature_weight,
                                   min_points, keep, out_host, None if own_copy_stream else copy_stream,
                                   chi2_table_len=max(need, 8 * int(chi2_table_len)))
    return out


_STAGING = {}
_IN_FLIGHT = []


def hold_until_copied(tensor, copy_stream):
    """Keep `tensor` alive until the copy just queued on `copy_stream` has read it.  (Tensor.record_stream does the
    same job but hands the block back to the caching allocator at an unpredictable time: in a tile stream the
    allocator then keeps growing with cudaMalloc calls of 10-300 ms in the middle of the stream.)  Finished copies
    are dropped here, so the allocator sees one free per column and reaches a steady state after two tiles."""
    done = torch.cuda.Event()
    done.record(copy_stream)
    _IN_FLIGHT[:] = [(e, t) for e, t in _IN_FLIGHT if not e.query()]
    _IN_FLIGHT.append((done, tensor))




def _staging_buffers(dev, nmax):
    """The two device staging buffers of the tile stream, kept between calls (a stream of equally sized tiles then
    allocates nothing); they are replaced when a larger tile arrives."""
    key = (dev.type, dev.index)
    cur = _STAGING.get(key)
    if cur is None or cur[0].shape[0] < nmax:
        _STAGING[key] = cur = [torch.empty((nmax, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    return cur


def release_copied():
    """Drop the tensors held for device->host copies that have finished (called when a stream of tiles has ended)."""
    _IN_FLIGHT[:] = [(e, t) for e, t in _IN_FLIGHT if not e.query()]


def release_staging_buffers():
    """Free the tile stream's cached staging buffers (2 x 24 B x the largest tile seen)."""
    _STAGING.clear()


def multiscale_saliency_tiles(host_tiles, radii=DEFAULT_RADII, out_hosts=None, device=None, keep_results=True,
                              **params):
    """The pipeline over a sequence of host-resident tiles (pinned (N,3) float64 tensors), tile after tile on one
    GPU, with the transfers taken off the critical path: tile k+1 is uploaded on a copy stream while tile k is
    computed, and the saliency columns of tile k travel back (into ``out_hosts[k]``, {radius: pinned float32 tensor})
    while tile k+1 is computed.  Returns the per-tile results of :func:`multiscale_saliency` (device tensors;
    ``keep_results=False`` returns None per tile so that a long stream does not hold every tile's columns in HBM --
    the results are then in ``out_hosts`` only); all copies have completed on return."""
    tiles = list(host_tiles)
    if not tiles:
        return []
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    cur = torch.cuda.current_stream(dev)
    up = torch.cuda.Stream(device=dev)
    down = torch.cuda.Stream(device=dev)
    tiles = [t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t, dtype=np.float64))
             for t in tiles]                     # numpy input: pageable memory, its copy is then synchronous
    # two device staging buffers, filled alternately: no allocation (and no allocator-induced synchronisation)
    # happens between the tiles, and a buffer is overwritten only after the kernels that read it were issued
    nmax = max(int(t.shape[0]) for t in tiles)
    bufs = _staging_buffers(dev, nmax)[:min(2, len(tiles))]
    freed = [None] * len(bufs)

    def upload(k):
        t, slot = tiles[k], k % len(bufs)
        with torch.cuda.stream(up):
            if freed[slot] is not None:
                up.wait_event(freed[slot])
            else:
                up.wait_stream(cur)              # first use: the buffer was allocated on the compute stream
            d = bufs[slot][:int(t.shape[0])]
            d.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(up)
        return d, ev

    results = []
    checks = []
    redo = []

    def settle(j):
        # the overflow word of tile j is read while a later tile computes; a tile whose chi2 table was too short (rare:
        # > 1023 active neighbours in one neighbourhood) is repeated at the end with the reported length
        need = checks[j]()
        checks[j] = None
        if need:
            redo.append((j, need))

    nxt = upload(0)
    for k in range(len(tiles)):
        d, ev = nxt
        if k + 1 < len(tiles):
            nxt = upload(k + 1)                # enqueued before tile k's kernels: overlaps them
        cur.wait_event(ev)
        pend = []
        res = multiscale_saliency(d, radii, out_host=None if out_hosts is None else out_hosts[k],
                                  copy_stream=down, deferred=pend, **params)
        checks.append(pend[0])
        results.append(res if keep_results else None)
        del res
        done = torch.cuda.Event()
        done.record(cur)
        freed[k % len(bufs)] = done
        if k > 0:
            settle(k - 1)
    settle(len(tiles) - 1)
    cur.wait_stream(down)
    cur.synchronize()
    release_copied()
    for j, need in redo:
        res = multiscale_saliency(tiles[j].to(dev), radii, out_host=None if out_hosts is None else out_hosts[j],
                                  chi2_table_len=need, **params)
        cur.synchronize()
        if keep_results:
            results[j] = res
    return results
