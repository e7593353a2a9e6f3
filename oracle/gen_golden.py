#!/usr/bin/env python3
"""Generate the golden fixtures under ``tests/golden/`` FROM THE LIVE, UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE.  Runs only in the authoring container (needs ``/root/reference``):

    python oracle/gen_golden.py

It puts ``oracle/_shim`` (the ``pykdtree`` -> scipy stand-in) and ``/root/reference`` on
``sys.path``, imports the reference package as shipped, runs it on seeded inputs and stores
inputs (or the seeds that regenerate them) together with the reference's outputs.  The GPU
box has no ``/root/reference``; the tests there replay these files.

Fixtures
--------
ref_fixture_100.npz   the reference tests' own fixture (tests/test_*.py: 100 pts, T = I):
                      vectorised AND loop-version H, g, e2 for the four classes.
linearize_10k.npz     C1 shape (10k unit-cube points, scale 1): H, g, e2 at T = I and at a
                      non-identity T for the four classes.
align_10k.npz         C1 shape, scale 0.1: final T, iteration count and per-iteration
                      (e2, |dx|) of ``align`` for the four classes, tol 1e-3 and 1e-6.
structures.npz        voxel build (mean/cov/norm/icov/keys), kNN normals (k=15 and k=5),
                      NN indices on a 20k-point synthetic slab.
b01_sub.npz           40k-point random subsample of data/B-01.pcd as target (stored), scan =
                      section-8d perturbation of it: final T of ``align`` for the four classes.
slab_200k.npz         200k-point synthetic slab (regenerated from its seed): H, g, e2 at the
                      initial T and final T for the four classes.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import point_cloud_registration as ref  # noqa: E402  (the live reference)
from point_cloud_registration_b200 import datasets as ds  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
T_NONID_SO3 = np.array([-0.05, 0.1, -0.2])
T_NONID_T = np.array([0.1, 0.0, -0.1])


def nonid_T():
    T = np.eye(4)
    T[:3, :3] = ref.expSO3(T_NONID_SO3)
    T[:3, 3] = T_NONID_T
    return T


def make(cls_name, **kw):
    return getattr(ref, cls_name)(**kw)


def traced_align(obj, source, init_T):
    """Run the reference's align() while recording what calc_H_g_e2 saw / returned."""
    rec = []
    orig = obj.calc_H_g_e2

    def spy(cur_T, src):
        H, g, e2 = orig(cur_T, src)
        rec.append((np.array(cur_T, dtype=np.float64), np.array(H), np.array(g), float(e2)))
        return H, g, e2

    obj.calc_H_g_e2 = spy
    try:
        T = obj.align(source, init_T=init_T)
    finally:
        obj.calc_H_g_e2 = orig
    return np.array(T), rec


CLASSES = (("ICP", {}), ("PlaneICP", {}), ("VPlaneICP", {}), ("NDT", {}))


def gen_ref_fixture():
    out = {}
    np.random.seed(42)
    target = np.random.rand(100, 3)
    R = ref.expSO3(np.array([0.1, 0.2, 0.3]))
    t = np.array([0.5, -0.3, 0.2])
    source = ((R @ target.T).T + t).astype(np.float32)
    out["target"], out["source"] = target, source
    for name, kw in (("ICP", {}), ("PlaneICP", {}), ("VPlaneICP", dict(voxel_size=1.0)),
                     ("NDT", dict(voxel_size=1.0))):
        o = make(name, max_iter=10, max_dist=2.0, tol=1e-3, **kw)
        o.set_target(target)
        if name == "PlaneICP":
            out["PlaneICP_normals"] = np.array(o.normal)
        H, g, e2 = o.calc_H_g_e2(np.eye(4), source)
        H2, g2, e22 = o.calc_H_g_e2_no_parallel_ver(np.eye(4), source)
        out[f"{name}_H"], out[f"{name}_g"], out[f"{name}_e2"] = H, g, e2
        out[f"{name}_H_loop"], out[f"{name}_g_loop"], out[f"{name}_e2_loop"] = H2, g2, e22
    np.savez_compressed(os.path.join(OUT, "ref_fixture_100.npz"), **out)


def gen_linearize_10k():
    out = {}
    target, source = ds.unit_cube_case(10000, scale=1.0, seed=42)
    source = source.astype(np.float32)
    out["n"], out["scale"], out["seed"] = 10000, 1.0, 42
    out["T_nonid"] = nonid_T()
    for name, kw in (("ICP", {}), ("PlaneICP", {}), ("VPlaneICP", dict(voxel_size=0.25)),
                     ("NDT", dict(voxel_size=0.25))):
        o = make(name, max_iter=30, max_dist=2.0, tol=1e-3, **kw)
        o.set_target(target)
        if name == "PlaneICP":
            out["PlaneICP_normals"] = np.array(o.normal)
        for tag, T in (("I", np.eye(4)), ("X", nonid_T())):
            H, g, e2 = o.calc_H_g_e2(T, source)
            out[f"{name}_{tag}_H"], out[f"{name}_{tag}_g"], out[f"{name}_{tag}_e2"] = H, g, e2
    np.savez_compressed(os.path.join(OUT, "linearize_10k.npz"), **out)


def gen_align_10k():
    out = {}
    target, source = ds.unit_cube_case(10000, scale=0.1, seed=42)
    out["n"], out["scale"], out["seed"] = 10000, 0.1, 42
    for name, kw in (("ICP", {}), ("PlaneICP", {}), ("VPlaneICP", dict(voxel_size=0.25)),
                     ("NDT", dict(voxel_size=0.25))):
        for tol in (1e-3, 1e-6):
            o = make(name, max_iter=30, max_dist=2.0, tol=tol, **kw)
            o.set_target(target)
            if name == "PlaneICP":
                out["PlaneICP_normals"] = np.array(o.normal)
            T, rec = traced_align(o, source, np.eye(4))
            tag = f"{name}_tol{tol:g}"
            out[f"{tag}_T"] = T
            out[f"{tag}_iters"] = len(rec)
            out[f"{tag}_e2"] = np.array([r[3] for r in rec])
            out[f"{tag}_Ts"] = np.array([r[0] for r in rec])
            out[f"{tag}_H0"], out[f"{tag}_g0"] = rec[0][1], rec[0][2]
    np.savez_compressed(os.path.join(OUT, "align_10k.npz"), **out)


def gen_structures():
    out = {}
    pts = ds.make_urban_slab(20000, seed=7)
    out["n"], out["seed"] = 20000, 7
    out["pts_checksum"] = np.float64(pts.astype(np.float64).sum())
    for vs in (0.5, 1.0):
        g = ref.VoxelGrid(vs)
        g.set_points(pts)
        g.calc_icov()
        tag = f"vox{vs:g}"
        out[f"{tag}_mean"], out[f"{tag}_cov"] = g.mean, g.cov
        out[f"{tag}_norm"], out[f"{tag}_icov"] = g.norm, g.icov
    out["keys_vs0.5"] = ref.voxel.get_keys(pts, 0.5)
    # also with float64 input and negative/positive mix (floor semantics)
    p64 = pts.astype(np.float64) * 3.7 - 11.0
    out["keys64_vs0.3"] = ref.voxel.get_keys(p64, 0.3)
    tree = ref.KDTree(pts)
    for k in (15, 5):
        out[f"normals_k{k}"] = ref.estimate_norm_with_tree(pts, tree, k)
    q = ds.perturb_scan(pts, seed=3)
    d, i = tree.query(q)
    out["nn_dist"], out["nn_idx"] = d, i.astype(np.int64)
    d, i = tree.query(pts[:2000], k=15)
    out["knn_dist"], out["knn_idx"] = d, i.astype(np.int64)
    out["voxel_filter_0.5"] = ref.voxel_filter(pts, 0.5)
    np.savez_compressed(os.path.join(OUT, "structures.npz"), **out)


def gen_b01_sub():
    out = {}
    full = ds.load_pcd_xyz(os.path.join(REF, "data", "B-01.pcd"))
    rng = np.random.default_rng(2025)
    target = full[np.sort(rng.choice(full.shape[0], 40000, replace=False))].copy()
    scan = ds.perturb_scan(target, seed=0)
    out["target"] = target
    out["scan_seed"] = 0
    out["scan_checksum"] = np.float64(scan.astype(np.float64).sum())
    tree = ref.KDTree(target)
    normals = ref.estimate_norm_with_tree(target, tree, 15)
    out["normals_k15"] = normals
    for name, kw in (("ICP", {}), ("PlaneICP", {}), ("VPlaneICP", dict(voxel_size=1.0)),
                     ("NDT", dict(voxel_size=2.0))):
        o = make(name, max_iter=30, max_dist=2.0, tol=1e-3, **kw)
        if name == "PlaneICP":
            o.set_target(target, tree, normals)
        else:
            o.set_target(target)
        T, rec = traced_align(o, scan, np.eye(4))
        out[f"{name}_T"], out[f"{name}_iters"] = T, len(rec)
        out[f"{name}_e2"] = np.array([r[3] for r in rec])
        out[f"{name}_H0"], out[f"{name}_g0"] = rec[0][1], rec[0][2]
    np.savez_compressed(os.path.join(OUT, "b01_sub.npz"), **out)


def gen_slab_200k():
    out = {}
    target = ds.make_urban_slab(200000, seed=11)
    scan = ds.perturb_scan(target, seed=5)
    out["n"], out["seed"], out["scan_seed"] = 200000, 11, 5
    out["target_checksum"] = np.float64(target.astype(np.float64).sum())
    out["scan_checksum"] = np.float64(scan.astype(np.float64).sum())
    tree = ref.KDTree(target)
    normals = ref.estimate_norm_with_tree(target, tree, 15)
    for name, kw in (("ICP", {}), ("PlaneICP", {}), ("VPlaneICP", dict(voxel_size=0.5)),
                     ("NDT", dict(voxel_size=1.0))):
        o = make(name, max_iter=30, max_dist=2.0, tol=1e-3, **kw)
        if name == "PlaneICP":
            o.set_target(target, tree, normals)
        else:
            o.set_target(target)
        T, rec = traced_align(o, scan, np.eye(4))
        out[f"{name}_T"], out[f"{name}_iters"] = T, len(rec)
        out[f"{name}_e2"] = np.array([r[3] for r in rec])
        out[f"{name}_H0"], out[f"{name}_g0"] = rec[0][1], rec[0][2]
        out[f"{name}_Hlast"], out[f"{name}_glast"], out[f"{name}_Tlast"] = rec[-1][1], rec[-1][2], rec[-1][0]
    np.savez_compressed(os.path.join(OUT, "slab_200k.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for fn in (gen_ref_fixture, gen_linearize_10k, gen_align_10k, gen_structures, gen_b01_sub,
               gen_slab_200k):
        print("generating", fn.__name__, flush=True)
        fn()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
