"""Point-to-plane ICP (reference point_cloud_registration/plane_icp.py:13-69)."""
import numpy as np

from . import _lib
from .kdtree import KDTree
from .registration import Registration


class PlaneICP(Registration):
    method = _lib.PLANE

    def __init__(self, max_iter=30, max_dist=2, tol=1e-3, k=15, device=None):
        super().__init__(max_iter=max_iter, tol=tol)
        self.max_dist = max_dist
        self.k = k
        self._device = device

    def set_target(self, target, kdree=None, norm=None):
        """Build the NN index and the k-NN normals on the GPU, or accept precomputed ones
        (plane_icp.py:19-28; the misspelt ``kdree`` keyword is the reference's)."""
        if _lib.is_device_array(target):
            self.target = _lib.as_f32_points(target, "target")     # stays on the GPU
        else:
            target = np.asarray(target)
            self.target = target.astype(np.float32)
        if isinstance(kdree, KDTree) and kdree.n == self.target.shape[0]:
            self.kdtree = kdree                      # reuse the caller's device-resident index
        else:
            self.kdtree = KDTree(self.target, device=self._device)
        self._ctx = self.kdtree._ctx
        self._ctx.build_correspondence_lists()           # shell lists streamed by the correspondence pass
        if kdree is None or norm is None:
            self._ctx.estimate_normals(self.k)
            self._normal = None
            self._given_normals = False
        else:
            nrm = _lib.as_f32_points(norm, "norm")
            if nrm.shape[0] != self.target.shape[0]:
                raise ValueError("norm must have one row per target point")
            self._ctx.set_normals(nrm)
            self._normal = norm
            self._given_normals = True
        self._target_ready()

    def update_target(self, target, norm=None):
        """Append ``target`` to the map (see Registration.update_target).  Normals: re-estimated over
        the enlarged map (k-NN neighbourhoods change where old and new points meet), unless the map was
        set with caller-supplied normals -- then ``norm`` must bring the normals of the new points."""
        if not self._is_target_set:
            raise ValueError("Target is not set.")
        had_own = self._normal is not None and self._given_normals
        if had_own and norm is None:
            raise ValueError("this map was set with caller-supplied normals: pass norm= for the new points")
        self.kdtree.append(target)
        self.target = self.kdtree.data
        self._ctx.build_correspondence_lists()
        if had_own:
            nrm = np.concatenate([np.asarray(self._normal, dtype=np.float32), np.asarray(norm, dtype=np.float32)])
            self._ctx.set_normals(np.ascontiguousarray(nrm))
            self._normal = nrm
        else:
            self._ctx.estimate_normals(self.k)
            self._normal = None

    @property
    def normal(self):
        """Per-target-point normals, caller's order (read back from the GPU on first use)."""
        if self._normal is None:
            self._normal = self._ctx.get_normals(self.target.shape[0])
        return self._normal
