"""Point-to-point ICP (reference point_cloud_registration/icp.py:12-57)."""
import numpy as np

from . import _lib
from .kdtree import KDTree
from .registration import Registration


class ICP(Registration):
    method = _lib.ICP

    def __init__(self, max_iter=30, max_dist=2, tol=1e-3, device=None):
        super().__init__(max_iter=max_iter, tol=tol)
        self.max_dist = max_dist
        self._device = device

    def set_target(self, target):
        """Upload the float32 target and build the exact-NN index on the GPU (icp.py:17-22)."""
        if not _lib.is_device_array(target):
            target = np.asarray(target).astype(np.float32)
        self.kdtree = KDTree(target, device=self._device)
        self.target = self.kdtree.data
        self._ctx = self.kdtree._ctx
        self._ctx.build_correspondence_lists()           # shell lists streamed by the correspondence pass
        self._target_ready()

    def update_target(self, target):
        """Append ``target`` to the map (see Registration.update_target)."""
        if not self._is_target_set:
            raise ValueError("Target is not set.")
        self.kdtree.append(target)
        self.target = self.kdtree.data
        self._ctx.build_correspondence_lists()
