"""CPU replay of the device functions (PCR_HD code in csrc/*.cuh) against the oracle.

tests/hostsim/_hostsim.so compiles the SAME headers the CUDA kernels use, for the host, so the
exact-search index arithmetic, the per-point algebra and the small linear algebra can be
verified without a GPU.  The GPU parity tests proper are tests/test_gpu_*.py."""
import ctypes as C
import os
import shutil
import sys

import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import pcr_oracle as orc
from point_cloud_registration_b200 import datasets as ds

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                                reason="nvcc not available to build the host replay")


@pytest.fixture(scope="module")
def hs():
    import build as hbuild
    lib = C.CDLL(hbuild.build())
    lib.hs_grid_build.restype = C.c_void_p
    lib.hs_grid_build.argtypes = [C.c_void_p, C.c_int64, C.c_double]
    lib.hs_grid_free.argtypes = [C.c_void_p]
    lib.hs_grid_cells.restype = C.c_int64
    lib.hs_grid_cells.argtypes = [C.c_void_p]
    lib.hs_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
    lib.hs_shell_build.restype = C.c_int64
    lib.hs_shell_build.argtypes = [C.c_void_p, C.c_double]
    lib.hs_shell_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.hs_nn_ball_first.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
    lib.hs_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    lib.hs_linearize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_gn_step.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_gn_step.restype = C.c_int
    lib.hs_so3_exp.argtypes = [C.c_void_p, C.c_void_p]
    lib.hs_solve6.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_solve6.restype = C.c_int
    lib.hs_eig3.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.hs_icov.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.hs_box_mask.restype = C.c_uint64
    lib.hs_box_mask.argtypes = [C.c_int] * 6
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def hs_nn(lib, pts, q, h, max_dist=1e18):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    g = lib.hs_grid_build(ptr(pts), len(pts), float(h))
    idx = np.empty(len(q), dtype=np.int64)
    dist = np.empty(len(q), dtype=np.float32)
    lib.hs_nn(g, ptr(q), len(q), float(max_dist), ptr(idx), ptr(dist))
    lib.hs_grid_free(g)
    return idx, dist


def test_box_mask(hs):
    rng = np.random.default_rng(0)
    for _ in range(300):
        lo = rng.integers(0, 4, 3)
        hi = np.array([rng.integers(l, 4) for l in lo])
        want = 0
        for z in range(lo[2], hi[2] + 1):
            for y in range(lo[1], hi[1] + 1):
                for x in range(lo[0], hi[0] + 1):
                    want |= 1 << (z * 16 + y * 4 + x)
        got = hs.hs_box_mask(int(lo[0]), int(hi[0]), int(lo[1]), int(hi[1]), int(lo[2]), int(hi[2]))
        assert got == want


def check_nn(lib, pts, q, h, max_dist=1e18):
    idx, dist = hs_nn(lib, pts, q, h, max_dist)
    d_ref, i_ref = orc.brute_force_knn(pts.astype(np.float32), q.astype(np.float32))
    inl = d_ref < max_dist * (1 - 1e-6)
    out = d_ref > max_dist * (1 + 1e-6)
    assert np.all(idx[out] == -1)
    assert np.all(idx[inl] >= 0)
    # same neighbour, or an exact tie in float32
    d_mine = np.linalg.norm(pts[idx[inl]].astype(np.float64) - q[inl].astype(np.float64), axis=1)
    assert np.all(np.abs(d_mine - d_ref[inl]) <= 2e-6 * np.maximum(1.0, d_ref[inl]))
    assert (idx[inl] == i_ref[inl]).mean() > 0.999
    assert np.allclose(dist[inl], d_ref[inl], rtol=1e-5, atol=1e-6)


def test_nn_unit_cube(hs):
    rng = np.random.default_rng(1)
    pts = rng.random((5000, 3)).astype(np.float32)
    q = (rng.random((3000, 3)) * 1.6 - 0.3).astype(np.float32)        # inside and outside the grid
    for h in (0.02, 0.07, 0.3, 2.0):
        check_nn(hs, pts, q, h)
    check_nn(hs, pts, q, 0.07, max_dist=0.05)
    check_nn(hs, pts, q, 0.07, max_dist=0.2)


def test_nn_slab(hs):
    pts = ds.make_urban_slab(20000, seed=7)
    q = ds.perturb_scan(pts, seed=3)
    check_nn(hs, pts, q[:4000], 0.2, max_dist=2.0)
    check_nn(hs, pts, q[:2000], 0.05, max_dist=2.0)
    far = q[:500] + np.array([0, 0, 5.0], dtype=np.float32)
    check_nn(hs, pts, far, 0.2, max_dist=2.0)
    check_nn(hs, pts, far, 0.2)
    z = load_golden("structures.npz")
    idx, _ = hs_nn(hs, pts, q, 0.2, 2.0)
    assert (idx == z["nn_idx"]).mean() > 0.9995


def test_nn_degenerate(hs):
    rng = np.random.default_rng(2)
    line = np.zeros((300, 3), dtype=np.float32)
    line[:, 0] = np.linspace(0, 10, 300)
    q = (rng.random((200, 3)) * np.array([12, 1, 1]) - np.array([1, 0.5, 0.5])).astype(np.float32)
    check_nn(hs, line, q, 0.1)
    one = np.array([[1.0, 2.0, 3.0]], dtype=np.float32)
    check_nn(hs, one, q, 0.5)
    dup = np.repeat(rng.random((50, 3)).astype(np.float32), 4, axis=0)
    idx, dist = hs_nn(hs, dup, dup, 0.1)
    assert np.all(dist == 0)
    nanq = np.array([[np.nan, 0, 0], [np.inf, 0, 0], [0.5, 0.5, 0.5]], dtype=np.float32)
    idx, _ = hs_nn(hs, dup, nanq, 0.1)
    assert idx[0] == -1 and idx[1] == -1 and idx[2] >= 0
    big = (rng.random((2000, 3)) * 800 - 400).astype(np.float32)      # large coordinates
    check_nn(hs, big, (rng.random((500, 3)) * 900 - 450).astype(np.float32), 25.0)


@pytest.mark.parametrize("dmax_frac", [1.0, 0.5, 1.7, 2.0, 2.6, 3.0])
def test_shell_lists_match_general_search(hs, dmax_frac):
    """shell_nn (margin-ordered per-cell lists, margin-bound termination, continuation into the
    general search) returns exactly what grid_search() returns, for every regime: on the surface,
    displaced by about one cell, far away (no list), outside the grid, NaN, tight and huge max_dist."""
    rng = np.random.default_rng(6)
    pts = ds.make_urban_slab(20000, seed=7)
    near = ds.perturb_scan(pts, seed=3)[:4001]
    off = near[:1500] + np.array([0.1, -0.2, 0.25], dtype=np.float32)
    far = near[:700] + np.array([0, 0, 3.0], dtype=np.float32)
    box = (rng.random((800, 3)) * (pts.max(0) - pts.min(0) + 6) + pts.min(0) - 3).astype(np.float32)
    bad = near[:40].copy()
    bad[::5, 1] = np.nan
    for h in (0.08, 0.3, 0.45, 1.5):
        g = hs.hs_grid_build(ptr(pts), len(pts), float(h))
        assert hs.hs_shell_build(g, dmax_frac) >= len(pts)
        for q, md, expect_lists in ((near, 2.0, h >= 0.45), (near, 0.03, h >= 0.45), (off, 2.0, h >= 1.5), (far, 2.0, False), (far, 1e9, False),
                                    (box, 2.0, False), (bad, 2.0, h >= 0.45), (pts[:2000], 1e30, True)):
            q = np.ascontiguousarray(q)
            i0 = np.empty(len(q), np.int64); d0 = np.empty(len(q), np.float32)
            hs.hs_nn(g, ptr(q), len(q), float(md), ptr(i0), ptr(d0))
            for pair in (0, 1):                               # shell_nn / the cursor API step by step
                i1 = np.empty(len(q), np.int64); d1 = np.empty(len(q), np.float32); used = np.zeros(len(q), np.uint8)
                hs.hs_shell_nn(g, ptr(q), len(q), float(md), ptr(i1), ptr(d1), ptr(used), pair)
                assert np.array_equal(d0, d1)
                assert np.array_equal(i0 < 0, i1 < 0) and (i0 == i1).mean() > 0.999
                if expect_lists:
                    assert used.mean() > 0.5
        hs.hs_grid_free(g)


def test_shell_lists_degenerate(hs):
    """Shell lists on degenerate targets: a line, one point, duplicates, huge coordinates, a target
    that sits in a single cell -- same answers as the brick-grid search."""
    rng = np.random.default_rng(12)
    line = np.zeros((300, 3), dtype=np.float32)
    line[:, 0] = np.linspace(0, 10, 300)
    one = np.array([[1.0, 2.0, 3.0]], dtype=np.float32)
    dup = np.repeat(rng.random((50, 3)).astype(np.float32), 4, axis=0)
    big = (rng.random((2000, 3)) * 800 - 400).astype(np.float32)
    blob = (rng.random((500, 3)) * 0.01 + 5.0).astype(np.float32)
    cases = ((line, 0.1, (rng.random((400, 3)) * np.array([12, 1, 1]) - np.array([1, 0.5, 0.5])).astype(np.float32)),
             (one, 0.5, (rng.random((200, 3)) * 4).astype(np.float32)),
             (dup, 0.1, np.concatenate([dup, dup + np.float32(0.03)])),
             (big, 25.0, (rng.random((500, 3)) * 900 - 450).astype(np.float32)),
             (blob, 1.0, (rng.random((300, 3)) * 3 + 3.5).astype(np.float32)))
    for pts, h, q in cases:
        q = np.ascontiguousarray(q)
        for frac in (2.0, 1.0):
            g = hs.hs_grid_build(ptr(pts), len(pts), float(h))
            assert hs.hs_shell_build(g, frac) >= len(pts)
            for md in (2.0, 1e9, 0.05):
                i0 = np.empty(len(q), np.int64); d0 = np.empty(len(q), np.float32)
                hs.hs_nn(g, ptr(q), len(q), float(md), ptr(i0), ptr(d0))
                i1 = np.empty(len(q), np.int64); d1 = np.empty(len(q), np.float32); used = np.zeros(len(q), np.uint8)
                hs.hs_shell_nn(g, ptr(q), len(q), float(md), ptr(i1), ptr(d1), ptr(used), 0)
                assert np.array_equal(d0, d1) and np.array_equal(i0 < 0, i1 < 0)
            hs.hs_grid_free(g)


def test_ball_first_search_matches_ring_growth(hs):
    """grid_search(ball_first=True) -- one pruned pass over the ball of max_dist, the kernels' path
    for list misses -- returns what the ring-by-ring search returns, for bounded and huge max_dist."""
    rng = np.random.default_rng(9)
    pts = ds.make_urban_slab(20000, seed=7)
    near = ds.perturb_scan(pts, seed=3)[:2000]
    far = near[:700] + np.array([0, 0, 1.7], dtype=np.float32)
    box = (rng.random((800, 3)) * (pts.max(0) - pts.min(0) + 6) + pts.min(0) - 3).astype(np.float32)
    for h in (0.1, 0.4, 1.5):
        g = hs.hs_grid_build(ptr(pts), len(pts), float(h))
        for q, md in ((near, 2.0), (far, 2.0), (far, 1.0), (box, 2.0), (box, 0.3), (far, 1e9)):
            q = np.ascontiguousarray(q)
            i0 = np.empty(len(q), np.int64); d0 = np.empty(len(q), np.float32)
            hs.hs_nn(g, ptr(q), len(q), float(md), ptr(i0), ptr(d0))
            i1 = np.empty(len(q), np.int64); d1 = np.empty(len(q), np.float32)
            hs.hs_nn_ball_first(g, ptr(q), len(q), float(md), ptr(i1), ptr(d1))
            assert np.array_equal(d0, d1) and np.array_equal(i0 < 0, i1 < 0) and (i0 == i1).mean() > 0.999
        hs.hs_grid_free(g)


def test_knn(hs):
    z = load_golden("structures.npz")
    pts = ds.make_urban_slab(20000, seed=7)
    g = hs.hs_grid_build(ptr(pts), len(pts), 0.2)
    for k in (15, 5, 1, 40):
        m = 2000
        idx = np.empty((m, k), dtype=np.int64)
        dist = np.empty((m, k), dtype=np.float32)
        q = np.ascontiguousarray(pts[:m])
        hs.hs_knn(g, ptr(q), m, k, ptr(idx), ptr(dist))
        d_ref, i_ref = orc.brute_force_knn(pts, q, k=k)
        d_ref = d_ref.reshape(m, k)
        assert np.allclose(dist, d_ref, rtol=1e-5, atol=1e-6)
        assert np.all(np.diff(dist, axis=1) >= 0)
        if k == 15:
            assert (idx == z["knn_idx"]).mean() > 0.999
    hs.hs_grid_free(g)
    tiny = np.random.default_rng(0).random((7, 3)).astype(np.float32)
    g = hs.hs_grid_build(ptr(tiny), 7, 0.3)
    idx = np.empty((7, 10), dtype=np.int64)
    dist = np.empty((7, 10), dtype=np.float32)
    hs.hs_knn(g, ptr(tiny), 7, 10, ptr(idx), ptr(dist))
    assert np.all(idx[:, 7:] == -1) and np.all(np.isinf(dist[:, 7:])) and np.all(idx[:, :7] >= 0)
    hs.hs_grid_free(g)


def T_nonid():
    T = np.eye(4)
    T[:3, :3] = orc.so3_exp(np.array([-0.05, 0.1, -0.2]))
    T[:3, 3] = [0.1, 0.0, -0.1]
    return T


@pytest.mark.parametrize("method", [orc.ICP, orc.PLANE, orc.VPLANE, orc.NDT])
@pytest.mark.parametrize("Tname", ["I", "X"])
def test_terms_against_oracle(hs, method, Tname):
    target = ds.make_urban_slab(20000, seed=21)
    scan = ds.perturb_scan(target, seed=4)
    T = np.eye(4) if Tname == "I" else T_nonid()
    tg = orc.build_target(method, target, max_dist=2.0, k=10, voxel_size=1.0)
    H, g, e2, n_in = orc.linearize(tg, T, scan)
    moved = orc.transform_scan_f32(T, scan)
    recs = np.zeros((len(scan), 9), dtype=np.float32)
    if method in (orc.ICP, orc.PLANE):
        dist, nn = tg.index.query(moved)
        recs[:, :3] = tg.points[nn]
        if method == orc.PLANE:
            recs[:, 3:6] = tg.normals[nn]
    else:
        dist, nn = orc.voxel_query(tg.voxels, moved)
        recs[:, :3] = tg.voxels.mean[nn]
        if method == orc.VPLANE:
            recs[:, 3:6] = tg.voxels.norm[nn]
        else:
            W = tg.voxels.icov[nn]
            recs[:, 3:9] = np.stack([W[:, 0, 0], W[:, 0, 1], W[:, 0, 2], W[:, 1, 1], W[:, 1, 2], W[:, 2, 2]], axis=1)
    ok = (dist < 2.0).astype(np.uint8)
    out = np.zeros(29)
    Tc = np.ascontiguousarray(T)
    hs.hs_linearize(method, ptr(Tc), ptr(scan), len(scan), ptr(recs), ptr(ok), ptr(out))
    Hm = np.zeros((6, 6))
    Hm[np.triu_indices(6)] = out[:21]
    Hm = Hm + np.triu(Hm, 1).T
    assert int(out[28]) == n_in
    assert rel_err(Hm, H) < 2e-5
    assert np.max(np.abs(out[21:27] - g)) < 2e-5 * max(np.max(np.abs(g)), np.sqrt(np.max(np.abs(H)) * e2))
    assert abs(out[27] - e2) < 2e-5 * e2


@pytest.mark.parametrize("method", [orc.ICP, orc.PLANE])
def test_host_replay_of_one_linearisation(hs, method):
    """The kernel pair replayed end to end on the host: transform -> shell-list correspondences
    (device code compiled for the CPU) -> per-point terms -> record; against the oracle's
    calc_H_g_e2 at a non-identity pose."""
    target = ds.make_urban_slab(20000, seed=21)
    scan = ds.perturb_scan(target, seed=4)
    T = T_nonid()
    tg = orc.build_target(method, target, max_dist=2.0, k=10)
    H, g, e2, n_in = orc.linearize(tg, T, scan)
    moved = np.ascontiguousarray(orc.transform_scan_f32(T, scan))
    gh = hs.hs_grid_build(ptr(target), len(target), 0.4)
    assert hs.hs_shell_build(gh, 2.0) >= len(target)
    idx = np.empty(len(moved), np.int64); dist = np.empty(len(moved), np.float32); used = np.zeros(len(moved), np.uint8)
    hs.hs_shell_nn(gh, ptr(moved), len(moved), 2.0, ptr(idx), ptr(dist), ptr(used), 0)
    hs.hs_grid_free(gh)
    ok = (idx >= 0).astype(np.uint8)
    nn = np.where(idx >= 0, idx, 0)
    recs = np.zeros((len(scan), 9), dtype=np.float32)
    recs[:, :3] = tg.points[nn]
    if method == orc.PLANE:
        recs[:, 3:6] = tg.normals[nn]
    out = np.zeros(29)
    Tc = np.ascontiguousarray(T)
    hs.hs_linearize(method, ptr(Tc), ptr(scan), len(scan), ptr(recs), ptr(ok), ptr(out))
    Hm = np.zeros((6, 6))
    Hm[np.triu_indices(6)] = out[:21]
    Hm = Hm + np.triu(Hm, 1).T
    assert int(out[28]) == n_in
    assert rel_err(Hm, H) < 2e-5
    assert np.max(np.abs(out[21:27] - g)) < 2e-5 * max(np.max(np.abs(g)), np.sqrt(np.max(np.abs(H)) * e2))
    assert abs(out[27] - e2) < 2e-5 * e2


def test_small_linalg(hs):
    rng = np.random.default_rng(5)
    for scale in (1.0, 0.1, 3e-3, 1e-3, 1e-5):
        w = rng.normal(size=3) * scale
        R = np.empty(9)
        hs.hs_so3_exp(ptr(w), ptr(R))
        assert np.allclose(R.reshape(3, 3), orc.so3_exp(w), rtol=0, atol=1e-15)
    for _ in range(20):
        A = rng.normal(size=(6, 6))
        H = A @ A.T + 1e-3 * np.eye(6)
        g = rng.normal(size=6)
        x = np.empty(6)
        assert hs.hs_solve6(ptr(np.ascontiguousarray(H)), ptr(g), ptr(x)) == 0
        assert np.allclose(x, np.linalg.solve(H, g), rtol=1e-9, atol=1e-12)
    assert hs.hs_solve6(ptr(np.zeros((6, 6))), ptr(np.ones(6)), ptr(np.empty(6))) == 1
    # Gauss-Newton step incl. stop-before-update
    z = load_golden("align_10k.npz")
    H0, g0 = z["ICP_tol0.001_H0"], z["ICP_tol0.001_g0"]
    rec = np.zeros(32)
    rec[:21] = H0[np.triu_indices(6)]
    rec[21:27] = g0
    T = np.eye(4)
    dx = np.empty(6)
    dxn = np.empty(1)
    assert hs.hs_gn_step(ptr(rec), 1e-3, ptr(T), ptr(dx), ptr(dxn)) == 0
    dx_ref = -np.linalg.solve(H0, g0)
    assert np.allclose(dx, dx_ref, rtol=1e-10) and np.allclose(T, orc.se3_plus(np.eye(4), dx_ref), atol=1e-14)
    assert np.allclose(T, z["ICP_tol0.001_Ts"][1], atol=1e-12)
    T2 = np.eye(4)
    assert hs.hs_gn_step(ptr(rec), 10.0, ptr(T2), ptr(dx), ptr(dxn)) == 1 and np.array_equal(T2, np.eye(4))
    # eigenvectors / inverse covariance against the golden voxel statistics of the live reference
    s = load_golden("structures.npz")
    cov = s["vox0.5_cov"]
    c6 = np.ascontiguousarray(np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], axis=1))
    v = np.empty((len(cov), 3))
    hs.hs_eig3(ptr(c6), len(cov), ptr(v))
    assert np.min(np.abs(np.einsum('ij,ij->i', v, s["vox0.5_norm"]))) > 1 - 1e-9
    ic = np.empty_like(cov)
    hs.hs_icov(ptr(np.ascontiguousarray(cov)), len(cov), ptr(ic))
    assert np.array_equal(ic, s["vox0.5_icov"])
