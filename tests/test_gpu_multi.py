"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): scan tile-sharded over 2 processes,
target replicated, 29-double NCCL all-reduce inside libpcr_b200 -- must reproduce the
single-GPU record and transform."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, name, kw, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import point_cloud_registration_b200 as pcr
        from point_cloud_registration_b200 import datasets as ds
        from point_cloud_registration_b200.distributed import attach
        target = ds.make_urban_slab(200_000, seed=21)
        scan = ds.perturb_scan(target, seed=22, num_points=150_001)
        reg = getattr(pcr, name)(max_iter=30, max_dist=2.0, tol=1e-3, device=rank, **kw)
        reg.set_target(target)
        attach(reg)
        T0 = np.eye(4)
        H, g, e2 = reg.calc_H_g_e2(T0, scan)            # full scan in, this rank's tile linearised, records all-reduced
        T = reg.align(scan, init_T=T0)
        iters = reg.last_iterations
        # a NEW target on the same object: the communicator must follow it (the scan stays sharded, so a
        # context without communicator would silently solve on half the scan)
        target2 = ds.make_urban_slab(180_000, seed=23)
        scan2 = ds.perturb_scan(target2, seed=24, num_points=120_003)
        reg.set_target(target2)
        H2, g2, e22 = reg.calc_H_g_e2(T0, scan2)
        T2 = reg.align(scan2, init_T=T0)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), H=H, g=g, e2=e2, T=T, iters=iters, H2=H2, e22=e22, T2=T2,
                 iters2=reg.last_iterations)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,kw", [("PlaneICP", dict(k=10)), ("NDT", dict(voxel_size=1.0))])
def test_two_gpu_matches_one_gpu(tmp_path, name, kw):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import point_cloud_registration_b200 as pcr
    from point_cloud_registration_b200 import datasets as ds
    mp.spawn(worker, args=(2, free_port(), name, kw, str(tmp_path)), nprocs=2, join=True)
    target = ds.make_urban_slab(200_000, seed=21)
    scan = ds.perturb_scan(target, seed=22, num_points=150_001)
    one = getattr(pcr, name)(max_iter=30, max_dist=2.0, tol=1e-3, **kw)
    one.set_target(target)
    H, g, e2 = one.calc_H_g_e2(np.eye(4), scan)
    T = one.align(scan, init_T=np.eye(4))
    it1 = one.last_iterations
    target2 = ds.make_urban_slab(180_000, seed=23)
    scan2 = ds.perturb_scan(target2, seed=24, num_points=120_003)
    one.set_target(target2)
    H2, g2, e22 = one.calc_H_g_e2(np.eye(4), scan2)
    T2 = one.align(scan2, init_T=np.eye(4))
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.max(np.abs(z["H"] - H)) < 1e-9 * np.max(np.abs(H))
        assert abs(float(z["e2"]) - e2) < 1e-9 * e2
        assert np.linalg.norm(z["T"] - T) < 1e-9
        assert int(z["iters"]) == it1
        assert np.max(np.abs(z["H2"] - H2)) < 1e-9 * np.max(np.abs(H2))
        assert abs(float(z["e22"]) - e22) < 1e-9 * e22
        assert np.linalg.norm(z["T2"] - T2) < 1e-9 and int(z["iters2"]) == one.last_iterations
