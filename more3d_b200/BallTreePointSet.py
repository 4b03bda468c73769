"""Drop-in for the reference's ``BallTreePointSet`` (Downsampling/BallTreePointSet.py:17-426).

Same constructor, methods, keyword spellings (``return_distnaces``) and return types; the spatial
index underneath is the device-resident uniform grid of ``libmore3d_b200.so`` instead of
``sklearn.neighbors.BallTree``.  Radius / kNN results are index sets bit-identical to the ball
tree's (compare as sorted sets: the reference's order inside a list is unspecified too).
"""
import numpy as np
import torch

from . import _lib
from .grid import UniformGrid, cell_for_radius


class BallTreePointSet(object):

    def __init__(self, points, leaf_size=40, **kwargs):
        # BallTreePointSet.py:34-42: numpy arrays (and open3d clouds, absent here) are accepted
        if isinstance(points, torch.Tensor):
            points = points.detach().cpu().numpy()
        if not isinstance(points, np.ndarray):
            print("Given type: " + str(type(points)) +
                  " as input. Points type should be numpy array or open3d PointCloud")
            raise ValueError("Wrong input object.")
        _lib.require_cuda()
        # sklearn coerces to float64 C-order (sklearn/neighbors/_binary_tree.pxi.tp:850)
        self.__points = np.ascontiguousarray(points, dtype=np.float64)
        if self.__points.ndim != 2 or self.__points.shape[1] != 3:
            raise ValueError("points must have shape (N, 3)")
        self.__leaf_size = int(leaf_size)
        self.__dev = torch.from_numpy(self.__points).cuda()
        self.__grids = {}
        self.__bounds = None
        self.__tree = None  # LOD-cell hierarchy, built on first use (see lodtree.py)

    # ---------------- internals -------------
    def _grid(self, cell):
        key = float(cell)
        g = self.__grids.get(key)
        if g is None:
            if len(self.__grids) >= 4:  # bound HBM held by cached grids
                _, old = self.__grids.popitem()
                old.close()
            g = UniformGrid(self.__dev, key)
            self.__grids[key] = g
        return g

    def _knn_cell(self, k):
        """Column width giving ~k points per 3x3 block for this cloud's mean planimetric density."""
        if self.__bounds is None:                       # the cloud is immutable: one reduction per object
            self.__bounds = _lib.cloud_bounds(self.__dev)
        b = self.__bounds
        ext = b[3:] - b[:3]
        area = max(float(ext[0]) * float(ext[1]), 1e-300)
        n = self.__points.shape[0]
        cell = np.sqrt(max(k, 8) * area / (4.0 * n))
        if not np.isfinite(cell) or cell <= 0:
            cell = 1.0
        return float(cell)

    def _hierarchy(self):
        if self.__tree is None:
            from .lodtree import LodTree
            self.__tree = LodTree(self.__dev, self.__leaf_size)
        return self.__tree

    @property
    def device_points(self):
        return self.__dev

    # ------------- PROPERTIES -----------------
    @property
    def numberOfNodes(self):
        return self._hierarchy().n_nodes

    @property
    def ballTreeLeaves(self):
        return self._hierarchy().leaves()

    @property
    def maxLevel(self):
        return self._hierarchy().max_level()

    @property
    def Size(self):
        return self.__points.shape[0]

    # ------------ GENERAL FUNCTIONS------------------
    def getNodeRadius(self, index):
        return self._hierarchy().radius(index)

    def getPointsOfNode(self, index):
        return self._hierarchy().points_of_node(index)

    def getSmallestNodesOfSize(self, radius, startingNode='root'):
        if startingNode == 'root':
            return self._hierarchy().smallest_nodes_of_size(radius)
        elif startingNode == 'leaves':
            # the reference's bottom-up mode is broken (functools.partial drops the mode,
            # BallTreePointSet.py:274): out of contract
            raise NotImplementedError("startingNode='leaves' is not executable in the reference")
        else:
            return None

    def getNodeLevel(self, index):
        return self._hierarchy().level(index)

    def getLeftChildOfNode(self, index):
        return self._hierarchy().left(index)

    def getRightChildOfNode(self, index):
        return self._hierarchy().right(index)

    def getParentOfNode(self, index):
        return self._hierarchy().parent(index)

    def getSiblingOfNode(self, index):
        return self._hierarchy().sibling(index)

    def query(self, pnts, k, return_distnaces=False):
        """k nearest neighbours, sorted by distance (BallTreePointSet.py:292-315)."""
        pnts = np.asarray(pnts, dtype=np.float64)
        if pnts.ndim == 1:
            pnts = pnts.reshape(1, -1)
        g = self._grid(self._knn_cell(k))
        idx, dist = g.knn(pnts, int(k), return_distance=bool(return_distnaces))
        indexes = idx.cpu().numpy()
        if return_distnaces:
            return indexes, dist.cpu().numpy()
        return indexes

    def queryRadius(self, pnts, radius):
        """Fixed-radius neighbours (BallTreePointSet.py:317-347): object array of int64 index
        arrays, one per query row; a single 1-D query point returns its index array directly."""
        if isinstance(pnts, list):
            pnts = np.array(pnts)
        single = False
        if isinstance(pnts, np.ndarray) and pnts.ndim == 1:
            pnts = pnts.reshape((1, -1))
            single = True
        pnts = np.ascontiguousarray(pnts, dtype=np.float64)
        radius = float(radius)
        g = self._grid(cell_for_radius(radius))
        self_query = (pnts.shape == self.__points.shape and
                      (pnts is self.__points or np.shares_memory(pnts, self.__points) or
                       np.array_equal(pnts, self.__points)))
        offsets, idx = g.radius_csr(None if self_query else pnts, radius, idx_dtype=torch.int64)
        off = offsets.cpu().numpy()
        flat = idx.cpu().numpy()
        out = np.empty(pnts.shape[0], dtype=object)
        o = off.tolist()                     # python ints: slicing with numpy scalars is 40 % slower
        rows = [flat[a:b] for a, b in zip(o[:-1], o[1:])]
        try:
            out[:] = rows
        except ValueError:                   # numpy versions that try to broadcast equal-length rows as a 2-D array
            for i, row in enumerate(rows):
                out[i] = row
        if single:
            return out[0]
        return out

    def queryRadiusCSR(self, pnts, radius):
        """Extension: the same query as device-resident CSR (offsets int64, indices int32), no host copy."""
        radius = float(radius)
        g = self._grid(cell_for_radius(radius))
        return g.radius_csr(pnts, radius)

    def ToNumpy(self):
        return self.__points

    def GetPoint(self, index):
        return self.ToNumpy()[index, :]
