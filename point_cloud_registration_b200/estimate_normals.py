"""k-NN normal estimation (reference point_cloud_registration/estimate_normals.py:11-105),
computed on the GPU: k nearest neighbours INCLUDING the point itself, float32 moment sums in
neighbour-rank order, smallest-eigenvalue eigenvector; the sign is arbitrary as in the
reference."""
import numpy as np

from .kdtree import KDTree


def estimate_norm_with_tree(points, kdtree, k=15):
    """estimate_normals.py:27-87.  ``kdtree`` should be a :class:`KDTree` built on ``points``
    (its device index is reused); any other object is ignored and a new index is built."""
    points = np.asarray(points)
    if not (isinstance(kdtree, KDTree) and kdtree.n == points.shape[0]):
        kdtree = KDTree(points)
    kdtree._ctx.estimate_normals(k)
    return kdtree._ctx.get_normals(points.shape[0])


def estimate_normals(points, k=15):
    """estimate_normals.py:11-24."""
    return estimate_norm_with_tree(points, KDTree(points), k=k)


def get_norm_lines(points, normals, length=0.1):
    """Line segments point -> point + length * normal for display (estimate_normals.py:91-105)."""
    lines = np.empty((2 * points.shape[0], points.shape[1]), dtype=points.dtype)
    lines[0::2] = points
    lines[1::2] = points + normals * length
    return lines
