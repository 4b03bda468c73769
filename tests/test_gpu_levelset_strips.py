"""GPU: the multi-GPU level set with G logical ranks (host threads) on one GPU == the single-GPU Solver."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RUN = dict(stepsize=.005, nu=1e-4, mu=2.5e-3, lambda1=1, lambda2=1, epsilon=1, cap=True, tolerance=0)


@pytest.mark.parametrize("G,init,model", [(2, "voxels", "chan_vese"), (3, "mask", "chan_vese"), (3, "voxels", "lmv"),
                                          (2, "mask", "lmv")])
def test_logical_strips_match_single_gpu(G, init, model):
    import torch
    from more3d_b200 import synthetic, levelsets_func as ls
    from more3d_b200.distributed_levelset import StripLevelSet, ThreadComm
    pts = synthetic.terrain(40000, seed=51, noise=0.03)
    pts -= pts.mean(axis=0)
    h, k = 0.5, 7
    rng = np.random.Generator(np.random.PCG64(2))
    zeta = np.column_stack([np.abs(np.sin(pts[:, 0] / 3.0)), rng.random(len(pts)) ** 2])
    if model == "lmv":
        zeta = zeta[:, :1] + 0.1 * zeta[:, 1:]          # the local-mean-and-variance model takes one channel
    nb, nrm, tang = ls.build_neighborhoods(pts.copy(), h, k)
    ref = ls.Solver(h, zeta, pts, nb, tang, model)
    mask = (pts[:, 0] ** 2 + pts[:, 1] ** 2) < 60.0
    if init == "voxels":
        ref.initialize_from_voxels(2.0)
        ref.phi[:] = ls.smooth(ref.phi, slice(None), nb, ref.diff.weights)
    else:
        ref.initialize_from_mask(mask)
    ref.run(12, **RUN)

    edges = np.linspace(pts[:, 0].min(), pts[:, 0].max(), G + 1)
    edges[0], edges[-1] = -np.inf, np.inf
    own = [np.nonzero((pts[:, 0] >= edges[r]) & (pts[:, 0] < edges[r + 1]))[0] for r in range(G)]
    hub = ThreadComm.Hub(G)
    out, err = [None] * G, []

    def worker(r):
        try:
            torch.cuda.set_device(0)
            s = StripLevelSet(ThreadComm(hub, r), pts[own[r]], zeta[own[r]], h, k, edges[r], edges[r + 1],
                              normals=nrm[own[r]], active_contour_model=model)
            if init == "voxels":
                s.initialize_from_voxels(2.0)
                s.smooth_phi()
            else:
                s.initialize_from_mask(mask[own[r]])
            s.run(12, **RUN)
            torch.cuda.synchronize()
            out[r] = (s.own_values(s.phi).cpu().numpy(), s.own_values(s.active).cpu().numpy().astype(bool), s.nloc - s.n,
                      list(s.rmsc_history))
        except Exception as e:      # pragma: no cover - surfaced below
            err.append((r, repr(e)))
            hub.barrier.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(G)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not err, err
    phi_ref, act_ref = ref.phi, ref.active
    for r in range(G):
        phi, act, ghosts, rms = out[r]
        assert ghosts > 0
        assert np.array_equal(act, act_ref[own[r]])                               # active sets: exact
        np.testing.assert_allclose(phi, phi_ref[own[r]], rtol=1e-9, atol=1e-11)   # region sums add in another order
        np.testing.assert_allclose(rms, ref.rmsc_history, rtol=1e-9)
