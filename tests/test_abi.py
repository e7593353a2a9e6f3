"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/pcr_b200.h declares; the Python layer fails loudly without a device."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pcr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from point_cloud_registration_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pcr_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes prototypes out of sync with the header"
    assert lib.pcr_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device every constructor must raise -- never compute on the CPU."""
    import ctypes
    from point_cloud_registration_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_int(0)
    has_gpu = lib.pcr_device_count(ctypes.byref(n)) == 0 and n.value > 0
    if has_gpu:
        pytest.skip("a GPU is present")
    import point_cloud_registration_b200 as pcr
    pts = np.random.default_rng(0).random((100, 3))
    for make in (lambda: pcr.ICP().set_target(pts), lambda: pcr.PlaneICP().set_target(pts),
                 lambda: pcr.VPlaneICP().set_target(pts), lambda: pcr.NDT().set_target(pts),
                 lambda: pcr.KDTree(pts), lambda: pcr.voxel_filter(pts, 0.5), lambda: pcr.estimate_normals(pts)):
        with pytest.raises(_lib.PcrError):
            make()
    with pytest.raises(ValueError, match="Target is not set."):
        pcr.ICP().align(pts)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "point_cloud_registration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_host_math_matches_reference_semantics():
    import point_cloud_registration_b200 as pcr
    from oracle import pcr_oracle as orc
    rng = np.random.default_rng(1)
    for s in (1.0, 1e-2, 3e-3, 1e-4):
        w = rng.normal(size=3) * s
        assert np.array_equal(pcr.expSO3(w), orc.so3_exp(w))
        T = pcr.makeT(pcr.expSO3(rng.normal(size=3)), rng.normal(size=3))
        dx = np.hstack([rng.normal(size=3), w])
        assert np.allclose(pcr.plus(T, dx), orc.se3_plus(T, dx), rtol=0, atol=1e-15)
    v = rng.normal(size=(40, 3)).astype(np.float32)
    assert np.allclose(pcr.skew2(v), orc.sum_hat_t_hat(v), rtol=1e-5)
    assert np.allclose(pcr.skews(v), orc.hat_batch(v))
    assert np.allclose(pcr.skew_time_vector(v, v[::-1]), orc.cross_rows(v, v[::-1]))
    R, t = pcr.makeRt(T)
    assert np.allclose(pcr.transform_points(T, v), v @ R.T + t)
    assert np.array_equal(pcr.get_keys(v * 10, 0.5), orc.voxel_keys(v * 10, 0.5))


def test_shard_bounds_cover_exactly():
    from point_cloud_registration_b200.distributed import shard_bounds
    for n in (0, 1, 7, 100, 1193011):
        for w in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_tiles_cover_exactly():
    """bench c5 / multi-GPU users: round-robin tiles partition the scan, each rank within tiles_per_rank points of
    the others."""
    from point_cloud_registration_b200.distributed import interleaved_tiles
    for n in (0, 1, 7, 1000, 100_000_003):
        for w in (1, 2, 3, 8):
            for t in (1, 4, 16):
                owned = [interleaved_tiles(n, r, w, t) for r in range(w)]
                flat = sorted(x for o in owned for x in o)
                assert sum(hi - lo for lo, hi in flat) == n
                assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
                if flat:
                    assert flat[0][0] == 0 and flat[-1][1] == n
                sizes = [sum(hi - lo for lo, hi in o) for o in owned]
                assert max(sizes) - min(sizes) <= t
    with pytest.raises(ValueError):
        interleaved_tiles(10, 2, 2)
