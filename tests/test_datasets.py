"""Workload generators (no GPU): the clouds the GPU path and the reference are compared on."""
import numpy as np
import pytest

from point_cloud_registration_b200 import datasets as ds
from point_cloud_registration_b200.voxel import get_keys


def test_generators_have_no_reference_key_collisions():
    """the reference partitions points by a LOSSY hash of the voxel coordinate
    (voxel.py:12-21, quirk Q9), this library by the exact coordinate.  The comparison workloads must not
    contain two voxels with one key (the generators assert it; this pins the assertion itself)."""
    b01 = ds.load_b01()
    if b01 is not None:
        ds.assert_no_key_collisions(b01, (0.5, 1.0, 2.0))
    ds.assert_no_key_collisions(ds.make_urban_slab(2_000_000, seed=10), (0.5, 1.0))
    with pytest.raises(AssertionError):                      # two voxels 10^10 apart in y share a key by construction
        ds.assert_no_key_collisions(np.array([[0.25, 0.25, 0.25], [0.25, 0.25 + 1e10, 0.25]]), (1.0,))




def test_reference_keys_match_the_package_helper():
    pts = np.random.default_rng(0).normal(0, 30, (5000, 3))
    assert np.array_equal(ds.reference_keys(pts, 0.5), get_keys(pts, 0.5))


def test_torch_and_numpy_collision_checks_agree():
    import torch
    pts = ds.make_urban_slab(200_000, seed=3)
    ds.assert_no_key_collisions(torch.from_numpy(pts), (0.5, 1.0))
    bad = torch.tensor([[0.25, 0.25, 0.25], [0.25, 0.25 + 1e10, 0.25]], dtype=torch.float64)
    with pytest.raises(AssertionError):
        ds.assert_no_key_collisions(bad, (1.0,))


def test_lever_arm_rotation_keeps_the_rim_displacement():
    so3 = (0.01, -0.02, 0.03)
    assert ds.lever_arm_so3(so3, 30.0) == so3                       # scenes no larger than C2 keep the section-8d rotation
    big = ds.lever_arm_so3(so3, 380.0)
    assert np.isclose(np.linalg.norm(big) * 380.0, np.linalg.norm(so3) * ds.C2_LEVER_ARM)


def test_b01_fixture_is_the_reference_cloud():
    b = ds.load_b01()
    if b is None:
        pytest.skip("data/b01_xyz.npz not present")
    assert b.shape == (ds.B01_POINTS, 3) and b.dtype == np.float32
    assert len(np.unique(b, axis=0)) == len(b)                       # no duplicate points: nearest neighbours are unique up to ties
