"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy restatement of the hot path of scomup/point-cloud-registration
(reference commit 5cedcb38): the per-iteration Gauss-Newton linearisation of
ICP / PlaneICP / VPlaneICP / NDT, the Gauss-Newton driver, and the once-per-target
builds (voxel statistics, closed-form inverse covariance, kNN normals).

Who may use this file: ``tests/``, ``__graft_entry__.smoke()`` (as the checker) and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package
``point_cloud_registration_b200`` never imports it and has no CPU fallback.

Pinning status
--------------
The reference ships NO golden vectors and never tests ``align()``:
its four tests only compare its vectorised ``calc_H_g_e2`` with its own loop version at
``T = I`` on 100 points.  This oracle is therefore pinned against OUTPUTS OF THE LIVE,
UNMODIFIED REFERENCE run in the authoring container (``oracle/gen_golden.py`` imports
``/root/reference`` with the ``pykdtree`` stand-in of ``oracle/_shim`` and writes
``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` replays those fixtures, and
``tests/test_oracle_vs_reference.py`` re-runs the comparison live whenever
``/root/reference`` exists.  The reference's own known-answer identity (vectorised ==
loop at T = I, atol 1e-3) is replayed in ``tests/test_oracle_golden.py`` too.

Third-party arithmetic
----------------------
The reference's nearest-neighbour search lives in ``pykdtree`` (un-vendored, un-pinned;
``setup.py:18-21``).  Its published contract -- exact Euclidean k-NN, result dtype follows
the tree dtype -- is restated by :class:`NNIndex` on top of ``scipy.spatial.cKDTree``
(the reference's own sanctioned alternative, ``kdtree.py:58-65``), and cross-checked by a
tree-free brute-force search (:func:`brute_force_knn`).

Every function cites the reference lines it follows.  Paths are relative to
``/root/reference/point_cloud_registration/``.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

SMALL_ANGLE_THETA2 = 1e-5          # math_tools.py:12  (epsilon)
HASH_MULT = 116101                 # voxel.py:17
HASH_MOD = 10000000000             # voxel.py:18

ICP, PLANE, VPLANE, NDT = 0, 1, 2, 3
METHOD_NAMES = {ICP: "icp", PLANE: "plane", VPLANE: "vplane", NDT: "ndt"}


# --------------------------------------------------------------------------------------
# SE(3) / SO(3) helpers                                     math_tools.py:61-113
# --------------------------------------------------------------------------------------
def hat(w):
    """3x3 cross-product matrix of a 3-vector (math_tools.py:61-64)."""
    wx, wy, wz = float(w[0]), float(w[1]), float(w[2])
    return np.array([[0.0, -wz, wy], [wz, 0.0, -wx], [-wy, wx, 0.0]])


def hat_batch(v):
    """(N,3) -> (N,3,3) float64 cross-product matrices (math_tools.py:34-41)."""
    out = np.zeros((v.shape[0], 3, 3))
    out[:, 0, 1] = -v[:, 2]
    out[:, 0, 2] = v[:, 1]
    out[:, 1, 0] = v[:, 2]
    out[:, 1, 2] = -v[:, 0]
    out[:, 2, 0] = -v[:, 1]
    out[:, 2, 1] = v[:, 0]
    return out


def cross_rows(a, b):
    """Row-wise a x b returned as float64, i.e. hat(a_i) @ b_i (math_tools.py:22-31)."""
    out = np.zeros((a.shape[0], 3))
    out[:, 0] = -a[:, 2] * b[:, 1] + a[:, 1] * b[:, 2]
    out[:, 1] = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    out[:, 2] = -a[:, 1] * b[:, 0] + a[:, 0] * b[:, 1]
    return out


def sum_hat_t_hat(v):
    """sum_i hat(v_i)^T hat(v_i) from six moment sums, in v's dtype (math_tools.py:44-58)."""
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    sxx, syy, szz = np.sum(x * x), np.sum(y * y), np.sum(z * z)
    sxy, sxz, syz = np.sum(x * y), np.sum(x * z), np.sum(y * z)
    return np.array([[szz + syy, -sxy, -sxz],
                     [-sxy, sxx + szz, -syz],
                     [-sxz, -syz, sxx + syy]])


def so3_exp(w):
    """Rodrigues formula with the reference's first-order branch for theta^2 <= 1e-5, which
    returns the NON-orthonormal I + hat(w) (math_tools.py:80-98, quirk Q10)."""
    w = np.asarray(w, dtype=np.float64)
    th2 = w.dot(w)
    W = hat(w)
    if th2 <= SMALL_ANGLE_THETA2:
        return np.eye(3) + W
    th = np.sqrt(th2)
    K = W / th
    return np.eye(3) + np.sin(th) * K + (1.0 - np.cos(th)) * K.dot(K)


def se3_plus(T, dx):
    """Right-multiplicative update T * [[Exp(dx[3:]), dx[:3]], [0, 1]] (math_tools.py:101-108)."""
    D = np.eye(4)
    D[:3, :3] = so3_exp(dx[3:])
    D[:3, 3] = dx[:3]
    return T @ D


def transform_scan_f32(T, scan_f32):
    """(R @ P^T)^T + t with T cast to float32 by the caller (math_tools.py:111-113 as called
    from icp.py:32, plane_icp.py:39, voxelized_plane_icp.py:32, ndt.py:26; quirk Q7)."""
    T32 = np.asarray(T).astype(np.float32)
    return (T32[:3, :3] @ scan_f32.T).T + T32[:3, 3]


# --------------------------------------------------------------------------------------
# Nearest neighbours                          kdtree.py:18-25 (pykdtree contract)
# --------------------------------------------------------------------------------------
class NNIndex:
    """Exact k-NN with pykdtree's dtype rules (see module docstring)."""

    def __init__(self, data):
        data = np.asarray(data)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        self.data = data
        self._tree = cKDTree(data)

    def query(self, pts, k=1):
        pts = np.asarray(pts)
        if self.data.dtype == np.float32 and pts.dtype != np.float32:
            raise TypeError("float32 index needs float32 queries")
        d, i = self._tree.query(pts, k=k, workers=-1)
        if self.data.dtype == np.float32:
            d = d.astype(np.float32)
        return d, i


def brute_force_knn(data, queries, k=1, chunk=2048):
    """Tree-free exact k-NN in float64 (validates NNIndex and the CUDA search)."""
    data = np.asarray(data, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    m = queries.shape[0]
    out_d = np.empty((m, k))
    out_i = np.empty((m, k), dtype=np.int64)
    d_sq = np.einsum('ij,ij->i', data, data)
    for s in range(0, m, chunk):
        q = queries[s:s + chunk]
        d2 = (np.einsum('ij,ij->i', q, q)[:, None] + d_sq[None, :] - 2.0 * q @ data.T)
        if k == 1:
            idx = np.argmin(d2, axis=1)[:, None]
        else:
            idx = np.argpartition(d2, k - 1, axis=1)[:, :k]
        # recompute the winners' distances exactly (the expansion above cancels badly)
        diff = q[:, None, :] - data[idx]
        dd = np.sqrt(np.einsum('mkj,mkj->mk', diff, diff))
        order = np.argsort(dd, axis=1, kind="stable")
        out_d[s:s + chunk] = np.take_along_axis(dd, order, axis=1)
        out_i[s:s + chunk] = np.take_along_axis(idx, order, axis=1)
    if k == 1:
        return out_d[:, 0], out_i[:, 0]
    return out_d, out_i


# --------------------------------------------------------------------------------------
# Voxel grid                                                   voxel.py:12-21, 69-179
# --------------------------------------------------------------------------------------
def voxel_coords(points, voxel_size):
    """floor(p / size) as int64 (voxel.py:16)."""
    return np.floor(points / voxel_size).astype(np.int64)


def voxel_keys(points, voxel_size=1.0):
    """The reference's lossy 64-bit polynomial hash of the voxel coordinate with Python
    floor-mod semantics (voxel.py:12-21, quirk Q9)."""
    c = voxel_coords(points, voxel_size)
    return (((c[:, 2] * HASH_MULT) % HASH_MOD + c[:, 1]) * HASH_MULT) % HASH_MOD + c[:, 0]


class VoxelStats:
    """Result of the reference's VoxelGrid.set_points (+ optional calc_icov)."""
    __slots__ = ("voxel_size", "min_points", "mean", "cov", "norm", "icov", "index", "count")


def voxel_build(points, voxel_size, min_points=10, with_icov=False):
    """Group by key, per-voxel mean, two-pass SAMPLE covariance (/max(n-1,1)), drop voxels
    with fewer than ``min_points`` points, normal = eigenvector of the smallest eigenvalue,
    NN index over the kept means (voxel.py:104-165; quirks Q2, Q3)."""
    points = np.asarray(points)
    keys = voxel_keys(points, voxel_size)
    _, inv = np.unique(keys, return_inverse=True)
    inv = inv.ravel()
    cnt = np.bincount(inv)
    mean = np.stack([np.bincount(inv, weights=points[:, a]) / cnt for a in range(3)], axis=1)
    dev = points - mean[inv]
    dx, dy, dz = dev[:, 0], dev[:, 1], dev[:, 2]
    denom = np.maximum(cnt - 1, 1)
    cxx = np.bincount(inv, weights=dx * dx) / denom
    cxy = np.bincount(inv, weights=dx * dy) / denom
    cxz = np.bincount(inv, weights=dx * dz) / denom
    cyy = np.bincount(inv, weights=dy * dy) / denom
    cyz = np.bincount(inv, weights=dy * dz) / denom
    czz = np.bincount(inv, weights=dz * dz) / denom
    cov = np.stack([np.stack([cxx, cxy, cxz], axis=1),
                    np.stack([cxy, cyy, cyz], axis=1),
                    np.stack([cxz, cyz, czz], axis=1)], axis=1)
    keep = cnt >= min_points
    vs = VoxelStats()
    vs.voxel_size, vs.min_points = voxel_size, min_points
    vs.mean, vs.cov, vs.count = mean[keep], cov[keep], cnt[keep]
    if vs.cov.shape[0]:
        _, vec = np.linalg.eigh(vs.cov)
        vs.norm = vec[:, :, 0]
    else:
        vs.norm = np.zeros((0, 3))
    vs.index = NNIndex(vs.mean)
    vs.icov = voxel_icov(vs.cov) if with_icov else None
    return vs


def voxel_icov(cov):
    """Adjugate / determinant inverse of each 3x3 covariance; det == 0 is replaced by 1e6,
    no regularisation (voxel.py:69-102, quirk Q6)."""
    a, b, c = cov[:, 0, 0], cov[:, 1, 1], cov[:, 2, 2]
    d, e, f = cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 2]
    det = a * (b * c) + 2 * (d * e) * f - a * (f * f) - b * (e * e) - c * (d * d)
    det = np.where(det == 0, 1000000.0, det)
    i00 = (b * c - f * f) / det
    i01 = -(d * c - e * f) / det
    i02 = (d * f - e * b) / det
    i11 = (a * c - e * e) / det
    i12 = -(a * f - d * e) / det
    i22 = (a * b - d * d) / det
    out = np.empty((cov.shape[0], 3, 3))
    out[:, 0, 0], out[:, 0, 1], out[:, 0, 2] = i00, i01, i02
    out[:, 1, 0], out[:, 1, 1], out[:, 1, 2] = i01, i11, i12
    out[:, 2, 0], out[:, 2, 1], out[:, 2, 2] = i02, i12, i22
    return out


def voxel_query(vs, pts):
    """1-NN of each query over the kept voxel MEANS (voxel.py:171-179, quirk Q2)."""
    return vs.index.query(pts)


def voxel_filter(points, voxel_size):
    """Per-voxel centroid down-sampling, float32 output, voxels in ascending key order
    (voxel.py:209-241)."""
    keys = voxel_keys(points, voxel_size)
    _, inv = np.unique(keys, return_inverse=True)
    inv = inv.ravel()
    cnt = np.bincount(inv).astype(np.float32)
    cnt[cnt == 0] = 1
    cols = [np.bincount(inv, weights=points[:, a]) / cnt for a in range(3)]
    return np.stack(cols, axis=1).astype(np.float32)


def color_by_voxel(points, voxel_size):
    """Packed RGB per point, one seeded random colour per voxel in ascending key order
    (voxel.py:183-206)."""
    keys = voxel_keys(points, voxel_size)
    uniq, inv = np.unique(keys, return_inverse=True)
    inv = inv.ravel()
    state = np.random.get_state()
    np.random.seed(42)
    colors = np.random.randint(0, 256, size=(len(uniq), 3), dtype=np.uint8)
    np.random.set_state(state)
    pc = colors[inv]
    rgb = pc[:, 0].astype(np.uint32) << 16 | pc[:, 1].astype(np.uint32) << 8 | pc[:, 2].astype(np.uint32)
    return np.rec.fromarrays([np.asarray(points).astype(np.float32), rgb], dtype=[('xyz', '<f4', (3,)), ('irgb', '<u4')])


# --------------------------------------------------------------------------------------
# kNN normals                                              estimate_normals.py:27-87
# --------------------------------------------------------------------------------------
def knn_moments_f32(points, nbr):
    """float32 sums of p and p p^T over the k neighbours IN RANK ORDER, then
    cov = E[pp^T] - mu mu^T in float32 (estimate_normals.py:41-72, quirk Q5)."""
    n, k = nbr.shape
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    prods = (x * x, y * y, z * z, x * y, x * z, y * z)
    s = np.zeros((n, 3), dtype=np.float32)
    m = np.zeros((6, n), dtype=np.float32)
    for r in range(k):
        j = nbr[:, r]
        s += points[j]
        for q in range(6):
            m[q] += prods[q][j]
    ppt = np.empty((n, 3, 3), dtype=np.float32)
    ppt[:, 0, 0], ppt[:, 1, 1], ppt[:, 2, 2] = m[0], m[1], m[2]
    ppt[:, 0, 1] = ppt[:, 1, 0] = m[3]
    ppt[:, 0, 2] = ppt[:, 2, 0] = m[4]
    ppt[:, 1, 2] = ppt[:, 2, 1] = m[5]
    mu = s / k
    return ppt / k - np.einsum('ij,ik->ijk', mu, mu)


def knn_normals(points, index, k=15):
    """Normals = eigenvector of the smallest eigenvalue of the k-NN covariance; k includes
    the point itself; sign arbitrary (estimate_normals.py:27-87)."""
    _, nbr = index.query(points, k=k)
    cov = knn_moments_f32(points, nbr.reshape(len(points), -1))
    _, vec = np.linalg.eigh(cov)
    return vec[:, :, 0]


# --------------------------------------------------------------------------------------
# Targets (what set_target builds)
# --------------------------------------------------------------------------------------
class Target:
    __slots__ = ("method", "points", "index", "normals", "voxels", "max_dist")


def build_target(method, target, *, max_dist=2.0, k=15, voxel_size=1.0, min_points=10,
                 index=None, normals=None):
    """icp.py:17-22, plane_icp.py:19-28, voxelized_plane_icp.py:18-21, ndt.py:18-22."""
    tg = Target()
    tg.method, tg.max_dist = method, max_dist
    tg.points = tg.index = tg.normals = tg.voxels = None
    if method == ICP:
        tg.points = np.asarray(target).astype(np.float32)
        tg.index = NNIndex(tg.points)
    elif method == PLANE:
        tg.points = np.asarray(target).astype(np.float32)
        if index is None or normals is None:
            tg.index = NNIndex(target)                      # tree keeps the caller's dtype
            tg.normals = knn_normals(np.asarray(target), tg.index, k)
        else:
            tg.index, tg.normals = index, normals
    elif method in (VPLANE, NDT):
        tg.voxels = voxel_build(target, voxel_size, min_points, with_icov=(method == NDT))
    else:
        raise ValueError("unknown method")
    return tg


# --------------------------------------------------------------------------------------
# Linearisation: (H, g, e2) = sum_i J_i^T W_i J_i, sum_i J_i^T W_i r_i, sum_i r_i^T W_i r_i
# --------------------------------------------------------------------------------------
def _assemble(H_tt, H_tr, H_rr, g_t, g_r):
    H = np.zeros((6, 6))
    H[:3, :3], H[:3, 3:], H[3:, :3], H[3:, 3:] = H_tt, H_tr, H_tr.T, H_rr
    return H, np.hstack([g_t, g_r])


def linearize_icp(tg, T, scan_f32):
    """icp.py:24-57.  r = R p + t - q;  J = [I, -R hat(p)]  BUT the rotational gradient is
    sum p x (R r) (``rs @ R.T``), not sum p x (R^T r)  -- quirk Q1, reproduced."""
    moved = transform_scan_f32(T, scan_f32)
    dist, nn = tg.index.query(moved)
    ok = dist < tg.max_dist                                       # strict, Euclidean (Q4)
    nn, moved, p = nn[ok], moved[ok], scan_f32[ok]
    n_in = moved.shape[0]
    r = moved - tg.points[nn]                                      # float32
    R = T[:3, :3]
    P = hat_batch(p)                                               # float64 (N,3,3)
    H_tr = -R @ hat(np.sum(p, axis=0))
    H, g = _assemble(n_in * np.eye(3), H_tr, sum_hat_t_hat(p),
                     r.sum(axis=0), np.einsum('nij,ni->j', P, -(r @ R.T)))
    return H, g, np.sum(r * r), n_in


def _plane_terms(R, p, moved, mu, nrm):
    """Shared algebra of plane_icp.py:46-67 and voxelized_plane_icp.py:41-62:
    r = n . (R p + t - mu);  J = [n^T, (p x R^T n)^T]."""
    r = np.einsum('ij,ij->i', nrm, moved - mu)
    Jt = nrm
    Jr = cross_rows(p, (R.T @ nrm.T).T)
    H, g = _assemble(np.einsum('ij,ik->jk', Jt, Jt), np.einsum('ij,ik->jk', Jt, Jr),
                     np.einsum('ij,ik->jk', Jr, Jr),
                     np.sum(Jt * r[:, None], axis=0), np.sum(Jr * r[:, None], axis=0))
    return H, g, np.sum(r * r)


def linearize_plane(tg, T, scan_f32):
    """plane_icp.py:30-69: correspondence = nearest target POINT and its normal."""
    moved = transform_scan_f32(T, scan_f32)
    dist, nn = tg.index.query(moved)
    ok = dist < tg.max_dist
    nn = nn[ok]
    H, g, e2 = _plane_terms(T[:3, :3], scan_f32[ok], moved[ok], tg.points[nn], tg.normals[nn])
    return H, g, e2, int(ok.sum())


def linearize_vplane(tg, T, scan_f32):
    """voxelized_plane_icp.py:23-64: correspondence = nearest kept voxel MEAN + its normal."""
    moved = transform_scan_f32(T, scan_f32)
    dist, nn = voxel_query(tg.voxels, moved)
    ok = dist < tg.max_dist
    nn = nn[ok]
    H, g, e2 = _plane_terms(T[:3, :3], scan_f32[ok], moved[ok], tg.voxels.mean[nn], tg.voxels.norm[nn])
    return H, g, e2, int(ok.sum())


def linearize_ndt(tg, T, scan_f32):
    """ndt.py:24-57: d = R p + t - mu; J = [I, -R hat(p)]; weight = Sigma^-1 of the nearest
    kept voxel mean."""
    moved = transform_scan_f32(T, scan_f32)
    dist, nn = voxel_query(tg.voxels, moved)
    ok = dist < tg.max_dist
    nn = nn[ok]
    W = tg.voxels.icov[nn]
    d = moved[ok] - tg.voxels.mean[nn]
    Jr = -T[:3, :3] @ hat_batch(scan_f32[ok])
    WJr = np.einsum('nij,njk->nik', W, Jr)
    Wd = np.einsum('nij,nj->ni', W, d)
    H, g = _assemble(np.sum(W, axis=0), np.sum(WJr, axis=0), np.einsum('nji,njk->ik', Jr, WJr),
                     np.sum(Wd, axis=0), np.einsum('nji,nj->i', Jr, Wd))
    return H, g, np.einsum('ni,ni->', d, Wd), int(ok.sum())


_LINEARIZE = {ICP: linearize_icp, PLANE: linearize_plane, VPLANE: linearize_vplane, NDT: linearize_ndt}


def linearize(tg, T, scan_f32):
    """Dispatch on the target's method -> (H (6,6), g (6,), e2, inlier count)."""
    return _LINEARIZE[tg.method](tg, np.asarray(T, dtype=np.float64), scan_f32)


# --------------------------------------------------------------------------------------
# Gauss-Newton driver                                         registration.py:71-113
# --------------------------------------------------------------------------------------
def gauss_newton(tg, scan, init_T=None, max_iter=30, tol=1e-3, trace=None):
    """Scan cast to float32 (Q7); solve H dx = -g; STOP BEFORE applying a dx with
    |dx| < tol (Q8); otherwise T <- T [+] dx.  A singular H raises LinAlgError (Q12)."""
    scan_f32 = np.asarray(scan).astype(np.float32)
    T = np.eye(4) if init_T is None else np.asarray(init_T, dtype=np.float64)
    for it in range(max_iter):
        H, g, e2, n_in = linearize(tg, T, scan_f32)
        dx = -np.linalg.solve(H, g)
        if trace is not None:
            trace.append(dict(it=it, T=T.copy(), H=H, g=g, e2=float(e2), n=n_in, dx=dx))
        if np.linalg.norm(dx) < tol:
            break
        T = se3_plus(T, dx)
    return T


# --------------------------------------------------------------------------------------
# Object-style facade with the reference's class surface (used by tests / bench so that
# parity tests read like the reference's own tests).
# --------------------------------------------------------------------------------------
class _OracleRegistration:
    method = None

    def __init__(self, max_iter=30, max_dist=2, tol=1e-3, **kw):
        self.max_iter, self.max_dist, self.tol, self.kw = max_iter, max_dist, tol, kw
        self.tg = None

    def is_target_set(self):
        return self.tg is not None

    def set_target(self, target, index=None, normals=None):
        self.tg = build_target(self.method, target, max_dist=self.max_dist, index=index,
                               normals=normals, **self.kw)

    def calc_H_g_e2(self, cur_T, source):
        H, g, e2, _ = linearize(self.tg, cur_T, source)
        return H, g, e2

    def align(self, source, init_T=None, trace=None):
        if self.tg is None:
            raise ValueError("Target is not set.")
        return gauss_newton(self.tg, source, init_T, self.max_iter, self.tol, trace)


class OracleICP(_OracleRegistration):
    method = ICP


class OraclePlaneICP(_OracleRegistration):
    method = PLANE

    def __init__(self, max_iter=30, max_dist=2, tol=1e-3, k=15):
        super().__init__(max_iter, max_dist, tol, k=k)


class OracleVPlaneICP(_OracleRegistration):
    method = VPLANE

    def __init__(self, voxel_size=1.0, max_iter=30, max_dist=2, tol=1e-3, min_points=10):
        super().__init__(max_iter, max_dist, tol, voxel_size=voxel_size, min_points=min_points)


class OracleNDT(_OracleRegistration):
    method = NDT

    def __init__(self, voxel_size=1.0, max_iter=30, max_dist=2, tol=1e-3, min_points=10):
        super().__init__(max_iter, max_dist, tol, voxel_size=voxel_size, min_points=min_points)
