"""Normal-distributions transform (reference point_cloud_registration/ndt.py:12-57)."""
from . import _lib
from .registration import Registration
from .voxel import VoxelGrid


class NDT(Registration):
    method = _lib.NDT

    def __init__(self, voxel_size=1.0, max_iter=30, max_dist=2, tol=1e-3, device=None):
        super().__init__(max_iter=max_iter, tol=tol)
        self.voxel_size = voxel_size
        self.max_dist = max_dist
        self._device = device

    def set_target(self, target):
        """Voxel statistics + closed-form inverse covariances on the GPU (ndt.py:18-22)."""
        self.voxels = VoxelGrid(self.voxel_size, device=self._device)
        self.voxels.set_points(target)
        self.voxels.calc_icov()
        self._ctx = self.voxels._ctx
        self._target_ready()

    def update_target(self, target):
        """Append ``target`` to the map and rebuild the voxel statistics (see Registration.update_target)."""
        if not self._is_target_set:
            raise ValueError("Target is not set.")
        self.voxels.add_points(target)
