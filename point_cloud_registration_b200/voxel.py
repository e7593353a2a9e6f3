"""Voxel grid with the reference's surface (reference point_cloud_registration/voxel.py:12-241):
``VoxelGrid(voxel_size, min_points=10).set_points(pts)``, ``.calc_icov()``, ``.query(pts, names)``,
attributes ``mean / cov / norm / icov / kdtree``; ``voxel_filter``; ``get_keys``.  Statistics are
built on the GPU (csrc/pcr_build.cu).  Voxels are grouped by their EXACT integer coordinate
``floor(p / voxel_size)`` -- the reference groups by a lossy hash of it (``get_keys``, quirk Q9),
which is identical whenever that hash has no collision on the data; voxel ORDER is the
library's own (the reference's, ascending hash key, is an artefact of ``np.unique``)."""
import numpy as np

from . import _lib


def get_keys(points, voxel_size=1.0):
    """The reference's 64-bit polynomial voxel hash with Python floor-mod semantics
    (voxel.py:12-21).  Host-side utility kept for API compatibility; the device build does
    not use it."""
    c = np.floor(np.asarray(points) / voxel_size).astype(np.int64)
    p, m = 116101, 10000000000
    return ((c[:, 2] * p % m + c[:, 1]) * p) % m + c[:, 0]


class _MeanIndex:
    """``VoxelGrid.kdtree``: nearest kept voxel mean (voxel.py:165,176)."""

    def __init__(self, ctx):
        self._ctx = ctx

    def query(self, pts, k=1):
        if k != 1:
            raise NotImplementedError("the voxel-mean index answers k=1 queries only")
        return self._ctx.voxel_query(_lib.as_f32_points(pts, "query points"))


class VoxelGrid:
    def __init__(self, voxel_size, min_points=10, device=None):
        self.voxel_size = voxel_size
        self.min_points = min_points
        self.kdtree = None
        self._device = device
        self._ctx = None
        self._cache = None
        self._host64 = None

    def set_points(self, points):
        """voxel.py:104-165 on the GPU (inverse covariances included: they cost nothing extra).
        float32 clouds stay resident on the GPU so that ``add_points`` can extend them in place."""
        if self._ctx is None:
            self._ctx = _lib.Context(self._device)
        self._host64 = None
        is_dev = _lib.is_device_array(points)
        if (is_dev and _lib.DevicePoints(points).dtype == np.float32) or (not is_dev and np.asarray(points).dtype != np.float64):
            self._ctx.set_target_points(_lib.as_f32_points(points, "points"))
            self._ctx.build_voxels_from_target(self.voxel_size, self.min_points, with_icov=True)
        else:
            # float64 input: the statistics are accumulated from the float64 values (as the reference does)
            self._ctx.build_voxels(points, self.voxel_size, self.min_points, with_icov=True)
            self._host64 = None if is_dev else np.asarray(points)
        self._cache = None
        self.kdtree = _MeanIndex(self._ctx)

    def add_points(self, points):
        """Extend the cloud and rebuild the voxel statistics: same result as
        ``set_points(concatenate(old, new))`` (per-voxel sums run in point order, old points first)."""
        if self._ctx is None:
            raise ValueError("set_points has not been called")
        if self._host64 is not None:
            self.set_points(np.concatenate([self._host64, np.asarray(points, dtype=np.float64)]))
            return
        self._ctx.append_target_points(_lib.as_f32_points(points, "points"))
        self._ctx.build_voxels_from_target(self.voxel_size, self.min_points, with_icov=True)
        self._cache = None

    def calc_icov(self):
        """Closed-form inverse covariances (voxel.py:69-102); already built by set_points."""
        if self._ctx is None:
            raise ValueError("set_points has not been called")

    def _fetch(self):
        if self._cache is None:
            mean, cov, norm, icov, count = self._ctx.get_voxels(with_icov=True)
            self._cache = dict(mean=mean, cov=cov, norm=norm, icov=icov, count=count)
        return self._cache

    mean = property(lambda self: self._fetch()["mean"])
    cov = property(lambda self: self._fetch()["cov"])
    norm = property(lambda self: self._fetch()["norm"])
    icov = property(lambda self: self._fetch()["icov"])
    count = property(lambda self: self._fetch()["count"])

    def query(self, points, names):
        """Nearest kept voxel per point + the requested attributes (voxel.py:171-179)."""
        dist, idx = self.kdtree.query(points)
        out = {name: getattr(self, name)[idx] for name in names}
        out['dist'] = dist
        return out


def voxel_filter(points, voxel_size, device=None):
    """Per-voxel centroid down-sampling, float32 (voxel.py:209-241), on the GPU."""
    return _lib.Context(device).voxel_filter(points, voxel_size)


def color_by_voxel(points, voxel_size, device=None):
    """Structured array ``[('xyz', '<f4', (3,)), ('irgb', '<u4')]`` colouring every point by its voxel
    (voxel.py:183-206).  Voxel membership comes from the GPU (exact integer coordinates); the
    colour table is the reference's (``np.random.seed(42)``, one colour per voxel in ascending order of
    the reference's hash key), so the result equals the reference's whenever that hash has no
    collision on the data (quirk Q9)."""
    pts = np.asarray(points)
    labels, coords = _lib.Context(device).voxel_labels(pts, voxel_size)
    c = coords.astype(np.int64)
    p, m = 116101, 10000000000
    keys = ((c[:, 2] * p % m + c[:, 1]) * p) % m + c[:, 0]            # get_keys on the voxel coordinates
    rank = np.empty(len(keys), dtype=np.int64)
    rank[np.argsort(keys, kind="stable")] = np.arange(len(keys))      # position of each voxel in np.unique(keys)
    state = np.random.get_state()
    np.random.seed(42)
    colors = np.random.randint(0, 256, size=(len(keys), 3), dtype=np.uint8)
    np.random.set_state(state)
    pc = colors[rank[labels]]
    rgb = pc[:, 0].astype(np.uint32) << 16 | pc[:, 1].astype(np.uint32) << 8 | pc[:, 2].astype(np.uint32)
    return np.rec.fromarrays([pts.astype(np.float32), rgb], dtype=[('xyz', '<f4', (3,)), ('irgb', '<u4')])
