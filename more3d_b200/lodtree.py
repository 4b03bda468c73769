"""Level-of-detail node hierarchy behind ``BallTreePointSet`` (ball-tree-equivalent, built on the GPU).

Mirrors what the reference keeps after ``BallTree.get_arrays()`` (Downsampling/BallTreePointSet.py:45-53):
``idx_array``, ``node_data`` (idx_start, idx_end, is_leaf, radius) and the parent / child / sibling /
level relations, which for sklearn's heap layout are index arithmetic (children of i = 2i+1, 2i+2)
instead of the reference's O(nodes^2) reconstruction (:57-126).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class LodTree:
    def __init__(self, xyz_dev, leaf_size=40):
        lib = _lib.load()
        self.n = int(xyz_dev.shape[0])
        self.leaf_size = int(leaf_size)
        nl, nn = C.c_int64(0), C.c_int64(0)
        check(lib.more3d_balltree_shape(self.n, self.leaf_size, C.byref(nl), C.byref(nn)))
        self.n_levels, self.n_nodes = int(nl.value), int(nn.value)
        dev = xyz_dev.device
        with torch.cuda.device(dev):
            self.idx_array_dev = torch.empty(self.n, dtype=torch.int64, device=dev)
            start = torch.empty(self.n_nodes, dtype=torch.int64, device=dev)
            end = torch.empty(self.n_nodes, dtype=torch.int64, device=dev)
            leaf = torch.empty(self.n_nodes, dtype=torch.int32, device=dev)
            rad = torch.empty(self.n_nodes, dtype=torch.float64, device=dev)
            check(lib.more3d_balltree_build(ptr(xyz_dev), self.n, self.leaf_size, ptr(self.idx_array_dev), ptr(start),
                                            ptr(end), ptr(leaf), ptr(rad), stream_ptr()))
        # node tables: the device copies drive the level-of-detail selection (smallest_nodes_dev); the host copies
        # behind the hierarchy getters are made on first use (2 M nodes at 50 M points: tens of ms of copies and numpy)
        self._start_dev, self._end_dev, self._leaf_dev, self._rad_dev = start, end, leaf, rad
        self._host = None
        self._idx_array = None

    def _tables(self):
        if self._host is None:
            h = {"idx_start": self._start_dev.cpu().numpy(), "idx_end": self._end_dev.cpu().numpy(),
                 "is_leaf": self._leaf_dev.cpu().numpy(), "node_radius": self._rad_dev.cpu().numpy()}
            # a node exists (was visited by the recursive build) iff it is the root or its parent is internal
            ids = np.arange(self.n_nodes)
            par = (ids - 1) // 2
            exists = np.ones(self.n_nodes, bool)
            for lev in range(1, self.n_levels):
                a, b = (1 << lev) - 1, min((1 << (lev + 1)) - 1, self.n_nodes)
                p = par[a:b]
                exists[a:b] = exists[p] & (h["is_leaf"][p] == 0)
            h["exists"] = exists
            h["levels"] = np.floor(np.log2(ids + 1)).astype(np.int64)
            self._host = h
        return self._host

    idx_start = property(lambda self: self._tables()["idx_start"])
    idx_end = property(lambda self: self._tables()["idx_end"])
    is_leaf = property(lambda self: self._tables()["is_leaf"])
    node_radius = property(lambda self: self._tables()["node_radius"])
    exists = property(lambda self: self._tables()["exists"])
    levels = property(lambda self: self._tables()["levels"])

    @property
    def idx_array(self):
        if self._idx_array is None:
            self._idx_array = self.idx_array_dev.cpu().numpy()
        return self._idx_array

    # -- reference getters (None where the reference's relation dictionaries have no entry) --------
    def _child(self, i, k):
        c = 2 * int(i) + k
        return c if (self.is_leaf[i] == 0 and c < self.n_nodes and self.exists[c]) else None

    def left(self, i):
        return self._child(i, 1)

    def right(self, i):
        return self._child(i, 2)

    def parent(self, i):
        return None if int(i) == 0 else (int(i) - 1) // 2

    def sibling(self, i):
        i = int(i)
        if i == 0:
            return None
        s = i + 1 if i % 2 == 1 else i - 1
        return s if (s < self.n_nodes and self.exists[s]) else None

    def level(self, i):
        return int(self.levels[i]) if self.exists[i] else 0

    def max_level(self):
        return int(self.levels[self.exists].max())

    def leaves(self):
        return np.nonzero(self.is_leaf == 1)[0]

    def radius(self, i):
        return self.node_radius[i]

    def points_of_node(self, i):
        return self.idx_array[self.idx_start[i]:self.idx_end[i]]

    def smallest_nodes_dev(self, radius):
        """getSmallestNodesOfSize on the device: (cells int64 [m], starts int64 [m + 1] = idx_start of every cell
        followed by n), both device tensors, cells in the reference's depth-first left-before-right order."""
        lib = _lib.load()
        dev = self._start_dev.device
        with torch.cuda.device(dev):
            cells = torch.empty(self.n_nodes, dtype=torch.int64, device=dev)
            starts = torch.empty(self.n_nodes + 1, dtype=torch.int64, device=dev)
            m = C.c_int64(0)
            check(lib.more3d_lod_select(ptr(self._start_dev), ptr(self._leaf_dev), ptr(self._rad_dev), self.n_nodes,
                                        self.n, float(radius), ptr(cells), ptr(starts), C.byref(m), stream_ptr()))
        m = int(m.value)
        return cells[:m], starts[:m + 1]

    def smallest_nodes_of_size(self, radius):
        """First node on every branch with radius < `radius` or leaf, in the reference's depth-first,
        left-before-right order (BallTreePointSet.py:128-147)."""
        sel = self.smallest_nodes_dev(radius)[0].cpu().numpy()
        if sel.size == 1:
            return int(sel[0])       # the reference returns a bare index when the root itself qualifies
        return sel
