#!/usr/bin/env python3
"""Extract the xyz columns of the reference's data/B-01.pcd into data/b01_xyz.npz (run in the
authoring container, where /root/reference exists; the npz travels with the repository)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from point_cloud_registration_b200 import datasets as ds  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/B-01.pcd"
xyz = ds.load_pcd_xyz(src)
assert xyz.shape == (ds.B01_POINTS, 3) and xyz.dtype == np.float32
np.savez_compressed(ds.B01_NPZ, xyz=xyz)
print(ds.B01_NPZ, os.path.getsize(ds.B01_NPZ), "bytes")
