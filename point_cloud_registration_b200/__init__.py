"""point_cloud_registration_b200 -- B200-native drop-in for the hot path of
scomup/point-cloud-registration: same class surface (``ICP / PlaneICP / VPlaneICP / NDT`` with
``set_target`` / ``align`` / ``calc_H_g_e2``, ``KDTree``, ``VoxelGrid``, ``estimate_normals`` ...),
every per-point computation in hand-written sm_100a CUDA kernels behind the C ABI of
``include/pcr_b200.h`` (reference exports: point_cloud_registration/__init__.py:1-10).

Importing the package does not need the GPU; constructing any object does (no CPU fallback)."""
from .registration import Registration, UploadedScan
from .math_tools import makeRt, expSO3, makeT, skews, skew, skew2, huber_weight, plus, transform_points, skew_time_vector
from .voxelized_plane_icp import VPlaneICP
from .plane_icp import PlaneICP
from .icp import ICP
from .ndt import NDT
from .kdtree import KDTree
from .voxel import VoxelGrid, voxel_filter, color_by_voxel, get_keys
from .estimate_normals import estimate_normals, get_norm_lines, estimate_norm_with_tree
from .caratheodory import caratheodory, fast_caratheodory, create_gn_set, gn_set_from_registration

__all__ = ["Registration", "UploadedScan", "ICP", "PlaneICP", "VPlaneICP", "NDT", "KDTree", "VoxelGrid",
           "voxel_filter", "color_by_voxel", "get_keys", "estimate_normals", "estimate_norm_with_tree", "get_norm_lines",
           "makeRt", "makeT", "expSO3", "plus", "skew", "skews", "skew2", "skew_time_vector",
           "transform_points", "huber_weight", "caratheodory", "fast_caratheodory", "create_gn_set", "gn_set_from_registration"]
