"""Voxelized point-to-plane ICP (reference point_cloud_registration/voxelized_plane_icp.py:12-64)."""
from . import _lib
from .registration import Registration
from .voxel import VoxelGrid


class VPlaneICP(Registration):
    method = _lib.VPLANE

    def __init__(self, voxel_size=1.0, max_iter=30, max_dist=2, tol=1e-3, device=None):
        super().__init__(max_iter=max_iter, tol=tol)
        self.voxel_size = voxel_size
        self.max_dist = max_dist
        self._device = device

    def set_target(self, target):
        """Voxel means / covariances / plane normals + NN index over the kept means, on the
        GPU (voxelized_plane_icp.py:18-21)."""
        self.voxels = VoxelGrid(self.voxel_size, device=self._device)
        self.voxels.set_points(target)
        self._ctx = self.voxels._ctx
        self._target_ready()

    def update_target(self, target):
        """Append ``target`` to the map and rebuild the voxel statistics (see Registration.update_target)."""
        if not self._is_target_set:
            raise ValueError("Target is not set.")
        self.voxels.add_points(target)
