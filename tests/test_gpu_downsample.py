"""GPU parity tests of the downsampling path: ball-tree-equivalent LOD hierarchy, saliency-driven cell
simplification, Otsu threshold and the voxel-grid baseline -- against the reference wrapper's golden
vectors (tests/golden/balltree_6k.npz), scikit-learn's BallTree and the oracle restatements."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def leaf_sets(idx_array, start, end, nodes):
    return [np.sort(idx_array[start[i]:end[i]]) for i in nodes]


def test_hierarchy_matches_reference_wrapper(golden_dir):
    from more3d_b200.BallTreePointSet import BallTreePointSet
    g = np.load(os.path.join(golden_dir, "balltree_6k.npz"))
    bt = BallTreePointSet(g["points"], leaf_size=40)
    n = int(g["numberOfNodes"])
    assert bt.numberOfNodes == n and bt.maxLevel == int(g["maxLevel"]) and bt.Size == len(g["points"])
    assert np.array_equal(bt.ballTreeLeaves, g["leaves"])
    t = bt._hierarchy()
    assert np.array_equal(t.idx_start, g["idx_start"]) and np.array_equal(t.idx_end, g["idx_end"])
    assert np.array_equal(t.is_leaf, g["is_leaf"])
    # node membership: identical point sets for every node (order inside a node is nth_element's)
    for i in range(n):
        assert np.array_equal(np.sort(bt.getPointsOfNode(i)), np.sort(g["idx_array"][g["idx_start"][i]:g["idx_end"][i]]))
    rad = np.array([bt.getNodeRadius(i) for i in range(n)])
    np.testing.assert_allclose(rad, g["radius"], rtol=1e-12, atol=0)
    for i in range(n):
        assert bt.getNodeLevel(i) == g["levels"][i]
        assert (bt.getLeftChildOfNode(i) if bt.getLeftChildOfNode(i) is not None else -1) == g["left"][i]
        assert (bt.getParentOfNode(i) if bt.getParentOfNode(i) is not None else -1) == g["parent"][i]
        assert (bt.getSiblingOfNode(i) if bt.getSiblingOfNode(i) is not None else -1) == g["sibling"][i]
    for lod in (1.0, 2.0, 5.0):
        assert np.array_equal(np.atleast_1d(bt.getSmallestNodesOfSize(lod)), g["cells_%g" % lod])
    assert bt.getSmallestNodesOfSize(1.0, startingNode="nonsense") is None
    assert np.array_equal(bt.GetPoint(17), g["points"][17]) and bt.ToNumpy().shape == g["points"].shape


@pytest.mark.parametrize("n,leaf,kind", [(200000, 40, "terrain"), (50000, 7, "cube"), (1000, 40, "dup"), (30, 40, "tiny")])
def test_hierarchy_matches_sklearn(n, leaf, kind):
    from more3d_b200 import synthetic
    from more3d_b200.BallTreePointSet import BallTreePointSet
    from oracle import ref_path as rp
    rng = np.random.Generator(np.random.PCG64(n))
    if kind == "terrain":
        pts = synthetic.terrain(n, seed=3)
    elif kind == "dup":      # many ties on every axis: the (value, index) comparator decides
        pts = rng.integers(0, 6, (n, 3)).astype(np.float64)
    else:
        pts = rng.random((n, 3)) * [3.0, 2.0, 5.0]
    bt = BallTreePointSet(pts, leaf_size=leaf)
    t = bt._hierarchy()
    tree = rp.build_tree(pts, leaf)
    _, idx_array, nodes, _ = tree.get_arrays()
    assert t.n_nodes == len(nodes)
    assert np.array_equal(t.idx_start, nodes["idx_start"]) and np.array_equal(t.idx_end, nodes["idx_end"])
    assert np.array_equal(t.is_leaf, nodes["is_leaf"])
    leaves = np.nonzero(nodes["is_leaf"] == 1)[0]
    ours = t.idx_array
    for i in leaves:
        a = np.sort(ours[t.idx_start[i]:t.idx_end[i]])
        b = np.sort(np.asarray(idx_array)[nodes["idx_start"][i]:nodes["idx_end"][i]])
        assert np.array_equal(a, b), i
    np.testing.assert_allclose(t.node_radius, nodes["radius"], rtol=1e-11, atol=1e-300)
    for lod in (0.5, 2.0):
        assert np.array_equal(np.atleast_1d(bt.getSmallestNodesOfSize(lod)), np.atleast_1d(rp.lod_cells(tree, lod)))


def test_voxel_downsample_matches_oracle():
    from more3d_b200 import synthetic
    from more3d_b200.Downsample import voxel_downsample
    from oracle import ref_path as rp
    pts = synthetic.terrain(60000, seed=6)
    for vs in (0.3, 1.0, 5.0):
        got, keys = voxel_downsample(pts, vs, return_keys=True)
        want, wkeys = rp.voxel_downsample(pts, vs)
        assert np.array_equal(keys, wkeys)              # voxel assignment: bit-exact
        assert np.array_equal(got, want)                # same members, same accumulation order -> same fp64 mean
    rng = np.random.Generator(np.random.PCG64(2))
    cube = rng.random((20000, 3)) * 7 - 3
    got, keys = voxel_downsample(cube, 0.5, return_keys=True)
    want, wkeys = rp.voxel_downsample(cube, 0.5)
    assert np.array_equal(keys, wkeys) and np.array_equal(got, want)
    one = voxel_downsample(np.array([[1.0, 2.0, 3.0]]), 1.0)
    assert np.array_equal(one, [[1.0, 2.0, 3.0]])


def test_otsu_matches_oracle():
    from more3d_b200.Downsample import threshold_otsu
    from oracle import ref_path as rp
    rng = np.random.Generator(np.random.PCG64(1))
    v32 = np.concatenate([rng.normal(0.2, 0.05, 50000), rng.normal(0.8, 0.1, 30000)]).astype(np.float32)
    assert threshold_otsu(v32) == rp.otsu_threshold(v32)
    v64 = rng.random(70001) ** 4
    assert threshold_otsu(v64) == rp.otsu_threshold(v64)
    assert threshold_otsu(np.full(10, 0.25)) == 0.25


@pytest.mark.parametrize("literal", [False, True])
def test_saliency_downsample_matches_oracle(literal):
    from more3d_b200 import synthetic
    from more3d_b200.Downsample import saliency_downsample
    from oracle import ref_path as rp
    pts = synthetic.terrain(40000, seed=9)
    rng = np.random.Generator(np.random.PCG64(5))
    # saliency proxy with spatial structure so some cells exceed the 60 % rule
    sal = (rng.random(len(pts)) ** 4 + 0.8 * (np.sin(pts[:, 0] / 3.0) > 0.7)).astype(np.float64)
    for lod in (1.0, 2.0, 5.0):
        out, det = saliency_downsample(pts, sal, lod, literal=literal, return_details=True)
        w_pts, w_flag, w_idx, w_cells = rp.saliency_downsample(pts, sal, lod, literal=literal)
        assert np.array_equal(det["cells"], w_cells)
        assert out.shape == (len(w_idx), 4)
        assert np.array_equal(out[:, 3], w_flag)
        # per cell: same kept set (order inside a cell is nth_element's), same representative
        start = 0
        for c in range(len(w_cells)):
            k = int((det["cell"] == c).sum())
            a, b = det["idx"][start:start + k], w_idx[start:start + k]
            fa = out[start:start + k, 3]
            assert np.array_equal(np.sort(a[fa == 1]), np.sort(b[fa == 1]))
            assert np.array_equal(a[fa == 0], b[fa == 0])
            start += k
        if not literal:
            assert 0 < det["keep_all"].sum() < len(w_cells)   # both branches exercised
