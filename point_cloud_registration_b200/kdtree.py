"""Exact nearest-neighbour facade with the contract the reference expects from
``pykdtree.kdtree.KDTree`` (reference point_cloud_registration/kdtree.py:18-25):
``KDTree(data)``, ``query(pts, k=1) -> (dist, idx)`` with Euclidean distances, ``(M,)`` results
for k = 1 and ascending ``(M, k)`` otherwise.  The index is a GPU-resident brick grid
(csrc/pcr_grid.cuh); coordinates are held in float32 on the device."""
import numpy as np

from . import _lib


class KDTree:
    def __init__(self, data, device=None):
        if _lib.is_device_array(data):
            self._dist_dtype = np.float32
        else:
            data = np.asarray(data)
            self._dist_dtype = np.float32 if data.dtype == np.float32 else np.float64
        self.data = _lib.as_f32_points(data, "data")
        self.n = self.data.shape[0]
        self._ctx = _lib.Context(device)
        self._ctx.set_target_points(self.data)
        self._ctx.build_nn_index()

    def append(self, points):
        """Add points to the indexed cloud (the old ones stay on the GPU) and rebuild the index: same
        result as ``KDTree(concatenate(old, new))``."""
        new = _lib.as_f32_points(points, "points")
        self._ctx.append_target_points(new)
        self._ctx.build_nn_index()
        if isinstance(self.data, np.ndarray) and isinstance(new, np.ndarray):
            self.data = np.concatenate([self.data, new])
        else:
            self.data = None                       # (partly) device resident: not mirrored on the host
        self.n += new.shape[0]

    def query(self, pts, k=1):
        q = _lib.as_f32_points(pts, "query points")
        dist, idx = self._ctx.knn(q, k)
        dist = dist.astype(self._dist_dtype, copy=False)
        idx = idx.astype(np.uint32) if self.n < 2**32 - 1 else idx.astype(np.uint64)
        if k == 1:
            return dist[:, 0], idx[:, 0]
        return dist, idx
