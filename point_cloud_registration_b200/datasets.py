"""Workload generators and a minimal PCD reader (host side, NumPy only).

The reference benchmarks on ``data/B-01.pcd`` through ``benchmark/test_data.py:21-44``
(scan = rigidly moved, noised copy of the map).  The xyz columns of that file travel with this
repository as ``data/b01_xyz.npz`` (CC BY 4.0, see data/README.md; made by tools/make_b01_npz.py)
and are what the C2 benchmark runs on; the 10M / 100M configurations are synthetic scenes with the
same character: 2-D surfaces embedded in 3-D at B-01's surface density
(~170 pts/m^2) so that kNN normals and voxel planes are meaningful and occupied 0.5 m voxels hold
>= 10 points.
"""
from __future__ import annotations

import numpy as np

B01_POINTS = 1_193_011            # size of the reference's data/B-01.pcd
SURFACE_DENSITY = 170.0           # points / m^2 measured on B-01


def rodrigues(w):
    """Plain Rodrigues rotation (always orthonormal) used only to build test scenes."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def make_urban_slab(n_points, seed=0, density=SURFACE_DENSITY, dtype=np.float32):
    """Synthetic "urban slab": undulating ground + randomly oriented vertical wall patches
    + box clutter, sampled uniformly at ``density`` points per square metre with 1 cm
    roughness.  Deterministic in (n_points, seed).  Returns (n_points, 3) ``dtype``."""
    rng = np.random.default_rng(seed)
    n = int(n_points)
    area = n / density
    n_ground = int(round(0.50 * n))
    n_wall = int(round(0.35 * n))
    n_box = n - n_ground - n_wall
    side = np.sqrt(0.50 * area)
    half = 0.5 * side

    def ground_z(x, y):
        return 0.30 * np.sin(x / 15.0) * np.cos(y / 11.0)

    out = np.empty((n, 3), dtype=np.float64)

    # ground
    gx = rng.uniform(-half, half, n_ground)
    gy = rng.uniform(-half, half, n_ground)
    out[:n_ground, 0], out[:n_ground, 1], out[:n_ground, 2] = gx, gy, ground_z(gx, gy)

    # walls: vertical rectangles, random yaw / size / position
    wall_area = 0.35 * area
    n_walls = max(8, int(round(wall_area / 69.0)))
    w_len = rng.uniform(5.0, 20.0, n_walls)
    w_hgt = rng.uniform(3.0, 8.0, n_walls)
    w_yaw = rng.uniform(0.0, np.pi, n_walls)
    w_cx = rng.uniform(-half, half, n_walls)
    w_cy = rng.uniform(-half, half, n_walls)
    cum = np.cumsum(w_len * w_hgt)
    which = np.searchsorted(cum, rng.uniform(0.0, cum[-1], n_wall), side="right")
    which = np.minimum(which, n_walls - 1)
    u = (rng.uniform(-0.5, 0.5, n_wall)) * w_len[which]
    v = rng.uniform(0.0, 1.0, n_wall) * w_hgt[which]
    wx = w_cx[which] + u * np.cos(w_yaw[which])
    wy = w_cy[which] + u * np.sin(w_yaw[which])
    sl = slice(n_ground, n_ground + n_wall)
    out[sl, 0], out[sl, 1], out[sl, 2] = wx, wy, ground_z(w_cx[which], w_cy[which]) + v

    # box clutter: points on the five visible faces of small axis-aligned boxes
    box_area = max(area - 0.50 * area - wall_area, 1.0)
    n_boxes = max(4, int(round(box_area / 14.0)))
    b_sz = rng.uniform(1.0, 3.0, (n_boxes, 3))
    b_c = np.stack([rng.uniform(-half, half, n_boxes), rng.uniform(-half, half, n_boxes)], axis=1)
    which = rng.integers(0, n_boxes, n_box)
    face = rng.integers(0, 5, n_box)                     # 0:+x 1:-x 2:+y 3:-y 4:top
    a = rng.uniform(-0.5, 0.5, n_box)
    b = rng.uniform(0.0, 1.0, n_box)
    sx, sy, sz = b_sz[which, 0], b_sz[which, 1], b_sz[which, 2]
    lx = np.where(face == 0, 0.5, np.where(face == 1, -0.5, a)) * sx
    ly = np.where(face == 2, 0.5, np.where(face == 3, -0.5, np.where(face == 4, rng.uniform(-0.5, 0.5, n_box), a))) * sy
    lz = np.where(face == 4, 1.0, b) * sz
    sl = slice(n_ground + n_wall, n)
    out[sl, 0] = b_c[which, 0] + lx
    out[sl, 1] = b_c[which, 1] + ly
    out[sl, 2] = ground_z(b_c[which, 0], b_c[which, 1]) + lz

    out += rng.normal(0.0, 0.01, out.shape)              # surface roughness
    rng.shuffle(out, axis=0)                             # acquisition order is not spatial
    return out.astype(dtype)


import os

B01_NPZ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "b01_xyz.npz")
C2_LEVER_ARM = 42.0      # metres: corner radius of the 1.19M-point slab (and ~ the half diagonal of B-01's busy part)


def load_b01():
    """(1193011, 3) float32 xyz of the reference's data/B-01.pcd from data/b01_xyz.npz, or None if
    the file is not there (callers then fall back to a synthetic slab of the same size and say so)."""
    if not os.path.exists(B01_NPZ):
        return None
    with np.load(B01_NPZ) as z:
        return np.ascontiguousarray(z["xyz"], dtype=np.float32)


def reference_keys(points, voxel_size):
    """The reference's lossy voxel hash (voxel.py:12-21, quirk Q9) -- same arithmetic as
    point_cloud_registration_b200.voxel.get_keys, repeated here so that the generators do not import
    the GPU package."""
    c = np.floor(np.asarray(points) / voxel_size).astype(np.int64)
    p, m = 116101, 10000000000
    return ((c[:, 2] * p % m + c[:, 1]) * p) % m + c[:, 0]


def assert_no_key_collisions(points, voxel_sizes=(0.5, 1.0)):
    """the reference groups points by a LOSSY hash of the voxel coordinate, this
    library by the exact coordinate.  The two partitions agree iff no two distinct voxels share a hash
    key -- assert it for every workload the two are compared on.  Works for NumPy arrays and torch
    tensors (the 100M-point clouds live on the GPU)."""
    for vs in voxel_sizes:
        if hasattr(points, "detach"):                      # torch tensor
            import torch
            c = torch.floor(points.to(torch.float32) / np.float32(vs)).to(torch.int64)
            p, m = 116101, 10000000000
            key = torch.remainder(torch.remainder(c[:, 2] * p, m) + c[:, 1], m)      # Python floor-mod semantics
            key = torch.remainder(key * p, m) + c[:, 0]
            lo = c.min(dim=0).values
            ext = (c.max(dim=0).values - lo + 1)
            exact = ((c[:, 2] - lo[2]) * ext[1] + (c[:, 1] - lo[1])) * ext[0] + (c[:, 0] - lo[0])
            n_key, n_exact = int(torch.unique(key).numel()), int(torch.unique(exact).numel())
        else:
            pts = np.asarray(points)
            c = np.floor(pts / pts.dtype.type(vs)).astype(np.int64)
            key = reference_keys(pts, pts.dtype.type(vs))
            lo = c.min(axis=0)
            ext = c.max(axis=0) - lo + 1
            exact = ((c[:, 2] - lo[2]) * ext[1] + (c[:, 1] - lo[1])) * ext[0] + (c[:, 0] - lo[0])
            n_key, n_exact = len(np.unique(key)), len(np.unique(exact))
        assert n_key == n_exact, (f"voxel size {vs}: the reference's get_keys hash merges {n_exact - n_key} voxels "
                                  "(quirk Q9): this cloud cannot be used to compare the two voxel partitions")


def lever_arm_so3(so3, points_radius, ref_radius=C2_LEVER_ARM):
    """Rotation vector scaled so that the scene's rim moves as far as C2's does: a fixed 0.037 rad
    displaces the corner of a 540 m slab by 14 m -- far beyond max_dist, where no ICP (the reference
    included) converges.  Scenes no larger than C2 keep the section-8d rotation."""
    return tuple(float(v) * min(1.0, ref_radius / max(float(points_radius), 1e-9)) for v in so3)


def perturb_scan(target, so3=(0.01, -0.02, 0.03), t=(0.1, -0.2, 0.3), sigma=0.005, seed=0,
                 num_points=None, dtype=np.float32):
    """scan = R * target + t (+ N(0, sigma^2)), optionally a random subsample without
    replacement -- the shape of the reference's ``generate_test_data``
    (benchmark/test_data.py:21-44) with the section-8d perturbation as default."""
    rng = np.random.default_rng(seed)
    pts = np.asarray(target, dtype=np.float64)
    if num_points is not None and num_points < pts.shape[0]:
        pts = pts[rng.choice(pts.shape[0], int(num_points), replace=False)]
    scan = pts @ rodrigues(so3).T + np.asarray(t, dtype=np.float64)
    if sigma > 0:
        scan = scan + rng.normal(0.0, sigma, scan.shape)
    return scan.astype(dtype)


def unit_cube_case(n=10000, scale=1.0, seed=42):
    """The reference tests' fixture shape (tests/test_icp.py:7-17) at size ``n``:
    target ~ U[0,1)^3 (float64), source = R target + t with so3 = scale*[.1,.2,.3],
    t = scale*[.5,-.3,.2]; ``scale=1`` reproduces the reference fixture exactly when n=100."""
    rs = np.random.RandomState(seed)
    target = rs.rand(n, 3)
    R = rodrigues(scale * np.array([0.1, 0.2, 0.3]))
    t = scale * np.array([0.5, -0.3, 0.2])
    source = (R @ target.T).T + t
    return target, source


def load_pcd_xyz(path):
    """Minimal reader for binary/ascii PCD files with float32 x,y,z as the first three
    fields (enough for the reference's data/B-01.pcd: 20-byte records
    ``<f4 x,y,z; u4 intensity; u4 rgb``).  Returns (N,3) float32."""
    with open(path, "rb") as fh:
        fields, sizes, counts, npts, mode = [], [], [], None, None
        while True:
            line = fh.readline()
            if not line:
                raise ValueError("PCD header ended without DATA line")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0].upper()
            if key == "FIELDS":
                fields = tok[1:]
            elif key == "SIZE":
                sizes = [int(v) for v in tok[1:]]
            elif key == "COUNT":
                counts = [int(v) for v in tok[1:]]
            elif key == "POINTS":
                npts = int(tok[1])
            elif key == "DATA":
                mode = tok[1].lower()
                break
        if fields[:3] != ["x", "y", "z"] or sizes[:3] != [4, 4, 4]:
            raise ValueError("unsupported PCD layout: need float32 x y z first")
        if not counts:
            counts = [1] * len(fields)
        stride = sum(s * c for s, c in zip(sizes, counts))
        if mode == "binary":
            raw = np.frombuffer(fh.read(npts * stride), dtype=np.uint8).reshape(npts, stride)
            return raw[:, :12].copy().view("<f4").reshape(npts, 3)
        if mode == "ascii":
            return np.loadtxt(fh, dtype=np.float32, usecols=(0, 1, 2)).reshape(-1, 3)
        raise ValueError(f"unsupported PCD DATA mode {mode!r}")


def make_urban_slab_torch(n_points, seed=0, density=SURFACE_DENSITY, device="cuda"):
    """Same scene family as :func:`make_urban_slab`, generated ON THE GPU with torch (used for
    the 10M / 100M benchmark configurations where host generation would take minutes).  Not
    bit-identical to the NumPy generator (different random streams).  Returns an (n,3)
    float32 tensor on ``device``; identical on every GPU for a given (n_points, seed)."""
    import math
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f64 = dict(dtype=torch.float64, device=device)

    def uni(lo, hi, size):
        return torch.rand(size, generator=g, **f64) * (hi - lo) + lo

    n = int(n_points)
    area = n / density
    n_ground = int(round(0.50 * n))
    n_wall = int(round(0.35 * n))
    n_box = n - n_ground - n_wall
    half = 0.5 * math.sqrt(0.50 * area)

    def ground_z(x, y):
        return 0.30 * torch.sin(x / 15.0) * torch.cos(y / 11.0)

    out = torch.empty((n, 3), **f64)
    gx, gy = uni(-half, half, n_ground), uni(-half, half, n_ground)
    out[:n_ground, 0], out[:n_ground, 1], out[:n_ground, 2] = gx, gy, ground_z(gx, gy)
    del gx, gy

    wall_area = 0.35 * area
    n_walls = max(8, int(round(wall_area / 69.0)))
    w_len, w_hgt = uni(5.0, 20.0, n_walls), uni(3.0, 8.0, n_walls)
    w_yaw = uni(0.0, math.pi, n_walls)
    w_cx, w_cy = uni(-half, half, n_walls), uni(-half, half, n_walls)
    cum = torch.cumsum(w_len * w_hgt, 0)
    which = torch.searchsorted(cum, uni(0.0, float(cum[-1]), n_wall), right=True).clamp_(max=n_walls - 1)
    u = uni(-0.5, 0.5, n_wall) * w_len[which]
    v = uni(0.0, 1.0, n_wall) * w_hgt[which]
    sl = slice(n_ground, n_ground + n_wall)
    out[sl, 0] = w_cx[which] + u * torch.cos(w_yaw[which])
    out[sl, 1] = w_cy[which] + u * torch.sin(w_yaw[which])
    out[sl, 2] = ground_z(w_cx[which], w_cy[which]) + v
    del which, u, v

    box_area = max(area - 0.50 * area - wall_area, 1.0)
    n_boxes = max(4, int(round(box_area / 14.0)))
    b_sz = uni(1.0, 3.0, (n_boxes, 3))
    b_cx, b_cy = uni(-half, half, n_boxes), uni(-half, half, n_boxes)
    which = torch.randint(0, n_boxes, (n_box,), generator=g, device=device)
    face = torch.randint(0, 5, (n_box,), generator=g, device=device)
    a, b, c = uni(-0.5, 0.5, n_box), uni(0.0, 1.0, n_box), uni(-0.5, 0.5, n_box)
    half_p, half_m = torch.full_like(a, 0.5), torch.full_like(a, -0.5)
    lx = torch.where(face == 0, half_p, torch.where(face == 1, half_m, a)) * b_sz[which, 0]
    ly = torch.where(face == 2, half_p, torch.where(face == 3, half_m, torch.where(face == 4, c, a))) * b_sz[which, 1]
    lz = torch.where(face == 4, torch.ones_like(b), b) * b_sz[which, 2]
    sl = slice(n_ground + n_wall, n)
    out[sl, 0] = b_cx[which] + lx
    out[sl, 1] = b_cy[which] + ly
    out[sl, 2] = ground_z(b_cx[which], b_cy[which]) + lz
    del which, face, a, b, c, lx, ly, lz

    out += torch.randn(out.shape, generator=g, **f64) * 0.01
    perm = torch.randperm(n, generator=g, device=device)
    return out[perm].to(torch.float32).contiguous()


def perturb_scan_torch(target, so3=(0.01, -0.02, 0.03), t=(0.1, -0.2, 0.3), sigma=0.005, seed=0):
    """Device version of :func:`perturb_scan` (full copy, no subsampling)."""
    import torch
    g = torch.Generator(device=target.device)
    g.manual_seed(int(seed) + 7919)
    R = torch.tensor(rodrigues(so3), dtype=torch.float64, device=target.device)
    tt = torch.tensor(np.asarray(t, dtype=np.float64), device=target.device)
    scan = target.to(torch.float64) @ R.T + tt
    if sigma > 0:
        scan += torch.randn(scan.shape, generator=g, dtype=torch.float64, device=target.device) * sigma
    return scan.to(torch.float32).contiguous()


def morton_order_torch(points, bits=10):
    """Permutation that orders an (n,3) CUDA tensor along a Morton curve of its own bounding box
    (2^bits cells per axis).  Multi-GPU runs shard the scan by CONTIGUOUS index ranges: sorted like
    this, every rank owns a spatial tile and streams only its part of the replicated
    target structure, instead of a random sample of the whole scene."""
    import torch
    lo = points.min(dim=0).values
    ext = (points.max(dim=0).values - lo).max().clamp_min(1e-9)
    g = ((points - lo) / ext * (2 ** bits - 1)).to(torch.int64).clamp_(0, 2 ** bits - 1)
    key = torch.zeros(points.shape[0], dtype=torch.int64, device=points.device)
    for b in range(bits):
        for a in range(3):
            key |= ((g[:, a] >> b) & 1) << (3 * b + a)
    return torch.argsort(key)

