"""Multi-GPU point-based level set: x-strips, one-hop ghost points, per-iteration halo exchanges and
all-reduces.  One process per GPU; the collectives go through a small communicator
interface (``DistComm`` = torch.distributed / NCCL; ``ThreadComm`` = G logical ranks as host threads on one GPU,
used by the tests).

Per Chan-Vese iteration (levelsets_func.py:598-642) the owner of a point computes it; what crosses ranks is
  * aux = (dp_, unit gradient) of the ghosts after phase 1          (32 B / ghost)
  * 4*C region sums (zeta_in / zeta_out, :410-412) and (sum inc^2, count)   all-reduce(SUM)
  * phi of the ghosts after the update, after the cap-smooth and after the lonely-smooth   (3 x 8 B / ghost;
    one more after the re-clamp of (de)activated points when active / last_active are separate masks)
  * activation flags of ghosts back to their owners (active[neighbors[active]] = True, :632)   (1 B / ghost, OR)
Ghost = a point of a neighbouring strip that appears in the k-neighbour list of an owned point.  The kNN
search itself needs a coordinate halo; its width is verified against the 2k-th neighbour distance.
"""
import ctypes as C
import queue
import threading

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .grid import UniformGrid


# ------------------------------------------------------------------------------------------------
class DistComm:
    """torch.distributed point-to-point + all-reduce (NCCL on GPUs)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange(self, send, recv_shapes, dtype, device):
        """send: {peer: tensor}; recv_shapes: {peer: shape}.  Returns {peer: tensor}."""
        dist = self.dist
        recv = {p: torch.empty(s, dtype=dtype, device=device) for p, s in recv_shapes.items()}
        ops = []
        for p, t in send.items():
            if t.numel():
                ops.append(dist.P2POp(dist.isend, t.contiguous(), p, self.group))
        for p, t in recv.items():
            if t.numel():
                ops.append(dist.P2POp(dist.irecv, t, p, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t


class ThreadComm:
    """G logical ranks as host threads of one process (tests: tile-and-halo on one GPU == untiled)."""

    class Hub:
        def __init__(self, world):
            self.world = world
            self.q = {(a, b): queue.Queue() for a in range(world) for b in range(world)}
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, hub, rank):
        self.hub, self.rank, self.world = hub, rank, hub.world

    def exchange(self, send, recv_shapes, dtype, device):
        for p, t in send.items():
            self.hub.q[(self.rank, p)].put(t.clone())
        out = {}
        for p, s in recv_shapes.items():
            out[p] = self.hub.q[(p, self.rank)].get(timeout=120).reshape(s)
        return out

    def _reduce(self, t, fn):
        h = self.hub
        h.slots[self.rank] = t.clone()
        h.barrier.wait(timeout=120)
        acc = h.slots[0].clone()
        for r in range(1, self.world):            # fixed order: identical result on every rank
            acc = fn(acc, h.slots[r])
        h.barrier.wait(timeout=120)
        t.copy_(acc)
        return t

    def allreduce_sum(self, t):
        return self._reduce(t, torch.add)

    def allreduce_max(self, t):
        return self._reduce(t, torch.maximum)


# ------------------------------------------------------------------------------------------------
class _State(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("phi", "zeta", "active", "last_active", "idx2", "laplace", "dphi", "aux",
                                          "bufA", "bufB", "azc", "outside", "part", "region", "incs")] + \
               [("n_own", C.c_int64)] + \
               [(n, C.c_double) for n in ("stepsize", "nu", "mu", "lambda1", "lambda2", "epsilon", "max_val")] + \
               [("channels", C.c_int), ("cap", C.c_int), ("model", C.c_int), ("reserved", C.c_int),
                ("stats", C.c_void_p), ("idxd", C.c_void_p), ("force", C.c_void_p)]


class StripLevelSet:
    """The Solver of one x-strip.  pts_own / zeta_own: this rank's points (n, 3) and cue (n, C) in any order;
    [x_lo, x_hi): the strip; all ranks call the constructor and ``run`` collectively."""

    def __init__(self, comm, pts_own, zeta_own, h, k, x_lo, x_hi, halo=None, normals=None,
                 active_contour_model='chan_vese'):
        lib = _lib.load()
        assert active_contour_model in ['chan_vese', 'lmv']
        self.active_contour_model = active_contour_model
        self.comm, self.h, self.k = comm, float(h), int(k)
        self.max_val = 2 * self.h
        dev = torch.device("cuda", torch.cuda.current_device())
        own = torch.as_tensor(pts_own, dtype=torch.float64).to(dev).contiguous()
        n = self.n = int(own.shape[0])
        zeta = torch.as_tensor(zeta_own, dtype=torch.float64).to(dev).reshape(n, -1).contiguous()
        # internal order of the owned points = grid-sorted (neighbour gathers become L1/L2 hits, see
        # levelsets_func.Derivative); `perm[i]` = caller's index of internal point i, `own_order()` maps back
        ext0 = (own.amax(0) - own.amin(0)).cpu().numpy()
        c0 = float(np.sqrt(8.0 * max(float(ext0[0]) * float(ext0[1]), 1e-300) / (4.0 * n)))
        g0 = UniformGrid(own, c0 if np.isfinite(c0) and c0 > 0 else 1.0)
        self.perm = g0.order().long()
        g0.close()
        self.inv = torch.empty_like(self.perm)
        self.inv[self.perm] = torch.arange(n, device=dev)
        own = own[self.perm].contiguous()
        zeta = zeta[self.perm].contiguous()
        if normals is not None:
            normals = torch.as_tensor(normals, dtype=torch.float64).to(dev)[self.perm]
        self.channels = int(zeta.shape[1])
        H = float(halo) if halo is not None else 6.0 * self.h
        left, right = comm.rank - 1, comm.rank + 1
        peers = [p for p in (left, right) if 0 <= p < comm.world]

        # 1. coordinate halo (kNN needs the neighbouring strips' boundary points)
        x = own[:, 0]
        sel = {p: torch.nonzero((x < x_lo + H) if p == left else (x >= x_hi - H)).flatten() for p in peers}
        cnt = comm.exchange({p: torch.tensor([sel[p].numel()], device=dev) for p in peers}, {p: (1,) for p in peers},
                            torch.int64, dev)
        halo_xyz = comm.exchange({p: own[sel[p]] for p in peers}, {p: (int(cnt[p]), 3) for p in peers}, torch.float64, dev)
        halo_rid = comm.exchange({p: sel[p] for p in peers}, {p: (int(cnt[p]),) for p in peers}, torch.int64, dev)
        cloud = torch.cat([own] + [halo_xyz[p] for p in peers]) if peers else own
        halo_owner = torch.cat([torch.full((int(cnt[p]),), p, dtype=torch.int64, device=dev) for p in peers]) if peers \
            else torch.zeros(0, dtype=torch.int64, device=dev)
        halo_remote = torch.cat([halo_rid[p] for p in peers]) if peers else torch.zeros(0, dtype=torch.int64, device=dev)

        # 2. 2k-NN of the owned points in the local cloud (build_neighborhoods, :158-163) + halo check
        ext = (cloud.amax(0) - cloud.amin(0)).cpu().numpy()
        cell = float(np.sqrt(max(2 * k, 8) * max(float(ext[0]) * float(ext[1]), 1e-300) / (4.0 * cloud.shape[0])))
        g = UniformGrid(cloud, cell if np.isfinite(cell) and cell > 0 else 1.0)
        knn, dist = g.knn(own, 2 * k, return_distance=True)
        g.close()
        reach = torch.minimum(x - x_lo if left >= 0 else torch.full_like(x, np.inf),
                              x_hi - x if right < comm.world else torch.full_like(x, np.inf)) + H
        ok = torch.tensor([float((dist[:, -1] < reach).all())], device=dev)
        comm.allreduce_sum(ok)
        if float(ok) < comm.world:
            raise ValueError("coordinate halo too narrow for the 2k-th neighbour distance: raise `halo`")

        # 3. frames of the owned points (compute_normals / compute_tangent_planes, :112-151)
        nrm = torch.empty((n, 3), dtype=torch.float64, device=dev)
        if normals is not None:
            nrm.copy_(torch.as_tensor(normals, dtype=torch.float64).to(dev))
        t1 = torch.empty((n, 3), dtype=torch.float64, device=dev)
        t2 = torch.empty((n, 3), dtype=torch.float64, device=dev)
        nb = torch.empty((n, k), dtype=torch.int32, device=dev)
        check(lib.more3d_ls_frames(ptr(cloud), n, k, ptr(knn), 1 if normals is not None else 0, ptr(nrm), ptr(t1), ptr(t2),
                                   ptr(nb), stream_ptr()))
        self.normals = nrm

        # 4. ghosts = halo points referenced by an owned point's k neighbours; compact cloud [own, ghosts]
        nbl = nb.long()
        ghost_local = torch.unique(nbl[nbl >= n])                       # sorted indices into `cloud`
        gcount = int(ghost_local.numel())
        remap = torch.full((int(cloud.shape[0]),), -1, dtype=torch.int64, device=dev)
        remap[:n] = torch.arange(n, device=dev)
        remap[ghost_local] = n + torch.arange(gcount, device=dev)
        nloc = self.nloc = n + gcount
        nb_all = torch.empty((nloc, k), dtype=torch.int32, device=dev)
        nb_all[:n] = remap[nbl].to(torch.int32)
        nb_all[n:] = torch.arange(n, nloc, device=dev, dtype=torch.int32)[:, None]   # ghost rows list themselves
        pts_all = torch.cat([own, cloud[ghost_local]])
        unit = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=dev)
        t1_all = torch.cat([t1, unit.expand(gcount, 3)]).contiguous()
        t2_all = torch.cat([t2, torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64, device=dev).expand(gcount, 3)]).contiguous()

        # 5. exchange lists: which of MY points each peer holds as ghosts, and where its values land here
        g_owner = halo_owner[ghost_local - n]
        g_remote = halo_remote[ghost_local - n]
        self.recv_idx = {p: (n + torch.nonzero(g_owner == p).flatten()) for p in peers}
        want = {p: g_remote[g_owner == p] for p in peers}
        wcnt = comm.exchange({p: torch.tensor([want[p].numel()], device=dev) for p in peers}, {p: (1,) for p in peers},
                             torch.int64, dev)
        self.send_idx = comm.exchange(want, {p: (int(wcnt[p]),) for p in peers}, torch.int64, dev)
        self.peers = peers

        # 6. WLS operator on the compact cloud (Derivative.__init__, :246-275)
        self.weights = torch.empty((nloc, k), dtype=torch.float64, device=dev)
        hnd, sing = C.c_void_p(), C.c_int(0)
        check(lib.more3d_ls_graph_create(ptr(pts_all), nloc, k, ptr(nb_all), ptr(t1_all), ptr(t2_all), 3 * self.h, 0,
                                         ptr(self.weights), None, C.byref(sing), stream_ptr(), C.byref(hnd)))
        self._g = hnd
        self.singular = bool(sing.value)
        self.points = pts_all
        self.neighbors = nb_all
        self.dev = dev

        # state
        f64 = dict(dtype=torch.float64, device=dev)
        self.phi = torch.zeros(nloc, **f64)
        self.zeta = torch.cat([zeta, torch.zeros((gcount, self.channels), **f64)]).contiguous()
        self.active = torch.zeros(nloc, dtype=torch.uint8, device=dev)
        self.last_active = torch.zeros(nloc, dtype=torch.uint8, device=dev)
        self.alias = False
        self._scratch = dict(
            idx2=torch.zeros(nloc, dtype=torch.uint8, device=dev), laplace=torch.zeros(nloc, **f64),
            dphi=torch.zeros(3 * nloc, **f64), aux=torch.zeros(4 * nloc, **f64), bufA=torch.zeros(nloc, **f64),
            bufB=torch.zeros(nloc, **f64), azc=torch.zeros(nloc, dtype=torch.uint8, device=dev),
            outside=torch.zeros(nloc, dtype=torch.uint8, device=dev),
            part=torch.zeros(int(lib.more3d_ls_scratch_doubles()), **f64), region=torch.zeros(4 * self.channels, **f64),
            incs=torch.zeros(2, **f64))
        self.rmsc_history = []
        if active_contour_model == 'lmv':
            # _lmv_step (levelsets_func.py:439-513): the data term reads the cue and the local statistics of every
            # neighbour, ghosts included -> the ghosts' zeta once, their statistics records every iteration.  Owners
            # always compute the record of a point some other rank holds as a ghost (`force`).
            if self.channels != 1:
                raise ValueError("operands could not be broadcast together: the 'lmv' model takes a "
                                 "single-channel cue, got %d channels" % self.channels)
            self._pull(self.zeta.view(-1))      # (nloc, 1) contiguous: the ghosts' cue, once
            self._stats = torch.zeros(4 * nloc, **f64)
            self._idxd = torch.zeros(nloc, dtype=torch.uint8, device=dev)
            self._force = torch.zeros(nloc, dtype=torch.uint8, device=dev)
            for p in peers:
                self._force[self.send_idx[p]] = 1

    def __del__(self):
        try:
            if getattr(self, "_g", None):
                _lib.load().more3d_ls_graph_destroy(self._g)
                self._g = None
        except Exception:
            pass

    # -- ghost traffic ----------------------------------------------------------------------------
    def _pull(self, t, width=1):
        """owners -> ghosts: t is a local array with `width` values per point (AoS)."""
        v = t.view(self.nloc, width) if width > 1 else t
        send = {p: v[self.send_idx[p]] for p in self.peers}
        shapes = {p: ((int(self.recv_idx[p].numel()), width) if width > 1 else (int(self.recv_idx[p].numel()),)) for p in self.peers}
        got = self.comm.exchange(send, shapes, t.dtype, self.dev)
        for p in self.peers:
            v[self.recv_idx[p]] = got[p]

    def _push_or(self, flags):
        """ghost activation flags -> their owners (OR), then cleared locally."""
        send = {p: flags[self.recv_idx[p]] for p in self.peers}
        shapes = {p: (int(self.send_idx[p].numel()),) for p in self.peers}
        got = self.comm.exchange(send, shapes, flags.dtype, self.dev)
        for p in self.peers:
            flags[self.send_idx[p]] |= got[p]
        flags[self.n:] = 0

    # -- initialisation (levelsets_func.py:547-596) --------------------------------------------------
    def _zero_crossings_all(self):
        lib = _lib.load()
        out = torch.empty(self.nloc, dtype=torch.uint8, device=self.dev)
        check(lib.more3d_ls_zero_crossings(self._g, ptr(self.phi), None, ptr(out), stream_ptr()))
        out[self.n:] = 0
        return out

    def initialize_from_voxels(self, h):
        lib = _lib.load()
        check(lib.more3d_ls_init_voxels(ptr(self.points), self.nloc, float(h), self.max_val, ptr(self.phi), stream_ptr()))
        self.active = self._zero_crossings_all()
        self.alias = True

    def own_values(self, t):
        """a local per-point tensor -> the owned points' values in the caller's original order"""
        return t[:self.n][self.inv]

    def initialize_from_mask(self, mask_own):
        m = torch.as_tensor(mask_own, dtype=torch.bool).to(self.dev)[self.perm]
        self.phi[:self.n] = torch.where(m, -self.max_val, self.max_val)
        self._pull(self.phi)
        self.active = self._zero_crossings_all()
        self.last_active = self.active.clone()
        self.alias = False

    def smooth_phi(self):
        """solver.phi[:] = smooth(solver.phi, slice(None), neighbors, weights) of the driver (run_levelsets.py:229)"""
        lib = _lib.load()
        out = torch.empty_like(self.phi)
        check(lib.more3d_ls_smooth(self._g, ptr(self.phi), None, ptr(out), stream_ptr()))
        self.phi[:self.n] = out[:self.n]
        self._pull(self.phi)

    # -- the iteration -----------------------------------------------------------------------------------
    def run(self, iterations, stepsize, nu, mu, lambda1, lambda2, epsilon, cap=True, tolerance=1e-4):
        lib = _lib.load()
        sc = self._scratch
        last = self.active if self.alias else self.last_active
        st = _State(phi=self.phi.data_ptr(), zeta=self.zeta.data_ptr(), active=self.active.data_ptr(),
                    last_active=last.data_ptr(), n_own=self.n, stepsize=stepsize, nu=nu, mu=mu, lambda1=lambda1,
                    lambda2=lambda2, epsilon=epsilon * self.h, max_val=self.max_val, channels=self.channels,
                    cap=1 if cap else 0, **{k: v.data_ptr() for k, v in sc.items()})
        lmv = self.active_contour_model == 'lmv'
        if lmv:
            st.model, st.stats, st.idxd, st.force = 1, self._stats.data_ptr(), self._idxd.data_ptr(), self._force.data_ptr()
        ph = lambda i: check(lib.more3d_ls_phase(self._g, i, C.byref(st), stream_ptr()))   # noqa: E731
        hist = torch.zeros((max(iterations, 1), 2), dtype=torch.float64, device=self.dev)
        done = 0
        for it in range(iterations):
            if not self.alias:
                ph(0)
                self._pull(self.phi)          # owners re-clamped some points: refresh their ghost copies
            ph(1)
            if nu > 0 or mu > 0 or lmv:
                self._pull(sc["aux"], 4)
            ph(2)
            if lmv:
                self._pull(self._stats, 4)     # the ghosts' local statistics from their owners
            else:
                self.comm.allreduce_sum(sc["region"])
            ph(3)
            self.comm.allreduce_sum(sc["incs"])
            self._pull(sc["bufA"])
            if not self.alias:
                self.last_active.copy_(self.active)
            ph(4)
            ph(5)
            self._pull(sc["bufA"])
            ph(6)
            self._pull(self.phi)
            ph(7)
            self._push_or(self.active)
            hist[it] = sc["incs"]
            done = it + 1
            if tolerance > 0:                    # the reference checks every iteration (:635-638): one small read-back
                inc = sc["incs"].cpu()
                if float((inc[0] / inc[1]) ** 0.5) < tolerance:
                    break
        h = hist[:done].cpu().numpy()
        with np.errstate(all="ignore"):
            r = np.sqrt(h[:, 0] / h[:, 1])
        self.rmsc_history.extend(float(v) for v in r)
        return bool(done and tolerance > 0 and r[-1] < tolerance)
