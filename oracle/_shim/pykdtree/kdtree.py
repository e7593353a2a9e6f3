"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party ``pykdtree.kdtree`` module.

The reference (scomup/point-cloud-registration) imports ``pykdtree.kdtree.KDTree``
(reference ``point_cloud_registration/kdtree.py:18-25``).  pykdtree is an un-vendored,
un-pinned third-party C/OpenMP library that is not installed in this image and cannot
be installed (no network).  The reference itself lists ``scipy.spatial.cKDTree`` as an
accepted backend (``kdtree.py:58-65``), so this stand-in delegates to it while keeping
the parts of pykdtree's documented contract the reference relies on:

* ``KDTree(data)``; ``query(pts, k=1) -> (dist, idx)``, Euclidean (not squared) distances;
* dtype follows the tree: a float32 tree requires float32 queries and returns float32
  distances; a float64 tree promotes the queries to float64;
* ``k > 1`` returns ``(M, k)`` arrays sorted by ascending distance.

It is put on ``sys.path`` only by ``oracle/gen_golden.py`` and by tests that import the
live reference from ``/root/reference`` in the authoring container.  Nothing in the
product package imports it.
"""
import numpy as np
from scipy.spatial import cKDTree


class KDTree:
    def __init__(self, data_pts, leafsize=16):
        data_pts = np.asarray(data_pts)
        if data_pts.dtype not in (np.float32, np.float64):
            data_pts = data_pts.astype(np.float64)
        self.data_pts = data_pts
        self.n = data_pts.shape[0]
        self._tree = cKDTree(data_pts, leafsize=leafsize)

    def query(self, query_pts, k=1, eps=0.0, distance_upper_bound=None, sqr_dists=False, mask=None):
        query_pts = np.asarray(query_pts)
        if self.data_pts.dtype == np.float32 and query_pts.dtype != np.float32:
            raise TypeError('Type mismatch. query points must be of type float32 '
                            'when data points are of type float32')
        ub = np.inf if distance_upper_bound is None else distance_upper_bound
        dist, idx = self._tree.query(query_pts, k=k, eps=eps, distance_upper_bound=ub, workers=-1)
        if sqr_dists:
            dist = dist * dist
        if self.data_pts.dtype == np.float32:
            dist = dist.astype(np.float32)
        return dist, idx
