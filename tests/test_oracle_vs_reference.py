"""Live comparison oracle <-> UNMODIFIED reference (only where /root/reference exists, i.e. in
the authoring container; skipped on the GPU box, where tests/golden/ stands in)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, frob, rel_err
from oracle import pcr_oracle as orc
from point_cloud_registration_b200 import datasets as ds

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "point_cloud_registration")),
                                reason="live reference not present on this machine")


@pytest.fixture(scope="module")
def ref():
    shim = os.path.join(ROOT, "oracle", "_shim")
    added = [p for p in (shim, REF) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    import point_cloud_registration as module
    yield module
    for p in added:
        sys.path.remove(p)


CASES = [("ICP", orc.OracleICP, {}), ("PlaneICP", orc.OraclePlaneICP, {}),
         ("VPlaneICP", orc.OracleVPlaneICP, dict(voxel_size=0.5)), ("NDT", orc.OracleNDT, dict(voxel_size=1.0))]


@pytest.mark.parametrize("name,ocls,kw", CASES)
def test_random_scene_live(ref, name, ocls, kw):
    target = ds.make_urban_slab(30000, seed=123)
    scan = ds.perturb_scan(target, so3=(0.02, 0.01, -0.03), t=(-0.15, 0.1, 0.2), seed=9)
    r = getattr(ref, name)(max_iter=30, max_dist=2.0, tol=1e-3, **kw)
    o = ocls(max_iter=30, max_dist=2.0, tol=1e-3, **kw)
    r.set_target(target)
    o.set_target(target)
    T0 = np.eye(4)
    T0[:3, :3] = ref.expSO3(np.array([0.01, 0.0, -0.01]))
    T0[:3, 3] = [0.05, 0.0, -0.02]
    Hr, gr, er = r.calc_H_g_e2(T0, scan)
    Ho, go, eo = o.calc_H_g_e2(T0, scan)
    assert rel_err(Ho, Hr) < 1e-6 and rel_err(go, gr) < 1e-5 and rel_err(eo, er) < 1e-6
    Tr = r.align(scan, init_T=T0)
    To = o.align(scan, T0)
    assert frob(To, Tr) < 1e-6


def test_math_tools_live(ref):
    rng = np.random.default_rng(0)
    for scale in (1.0, 1e-2, 1e-3, 1e-4):
        w = rng.normal(size=3) * scale
        assert np.array_equal(orc.so3_exp(w), ref.expSO3(w))
        dx = np.hstack([rng.normal(size=3), w])
        T = ref.makeT(ref.expSO3(rng.normal(size=3)), rng.normal(size=3))
        assert np.allclose(orc.se3_plus(T, dx), ref.plus(T, dx), rtol=0, atol=1e-15)
