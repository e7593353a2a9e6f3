"""Multi-process (world_size 2, gloo, CPU) tests of the scan-sharding logic of SURVEY.md 8e:
every rank linearises its own contiguous tile of the scan against the replicated target (the
oracle stands in for the GPU kernel here), the 29-double records are summed with an all-reduce,
and every rank must end up with the record -- and hence the Gauss-Newton step -- of the full scan."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pcr_oracle as orc
from point_cloud_registration_b200 import datasets as ds
from point_cloud_registration_b200.distributed import allreduce_record_host, exchange_unique_id, interleaved_tiles, shard_bounds


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def record_of(H, g, e2, n):
    rec = np.zeros(29)
    rec[:21] = H[np.triu_indices(6)]
    rec[21:27] = g
    rec[27], rec[28] = e2, n
    return rec


def worker(rank, world, port, method, out_dir, tiles_per_rank=0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        target = ds.make_urban_slab(30000, seed=3)
        scan = ds.perturb_scan(target, seed=4, num_points=20001)      # odd size: uneven tiles
        tg = orc.build_target(method, target, max_dist=2.0, k=10, voxel_size=1.0)
        T = np.eye(4)
        T[:3, 3] = [0.02, -0.01, 0.03]
        lo, hi = shard_bounds(len(scan), rank, world)
        mine = scan[lo:hi]
        if tiles_per_rank:                                            # bench c5: spatial tiles dealt round-robin
            mine = np.concatenate([scan[a:b] for a, b in interleaved_tiles(len(scan), rank, world, tiles_per_rank)])
        rec = allreduce_record_host(record_of(*orc.linearize(tg, T, mine)))
        uid = exchange_unique_id(lambda: b"x" * 128, rank)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), rec=rec, lo=lo, hi=hi, uid=np.frombuffer(uid, dtype=np.uint8))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("method", [orc.ICP, orc.PLANE, orc.NDT])
def test_sharded_linearisation_matches_full(tmp_path, method):
    world = 2
    mp.spawn(worker, args=(world, free_port(), method, str(tmp_path)), nprocs=world, join=True)
    target = ds.make_urban_slab(30000, seed=3)
    scan = ds.perturb_scan(target, seed=4, num_points=20001)
    tg = orc.build_target(method, target, max_dist=2.0, k=10, voxel_size=1.0)
    T = np.eye(4)
    T[:3, 3] = [0.02, -0.01, 0.03]
    full = record_of(*orc.linearize(tg, T, scan))
    recs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert recs[0]["lo"] == 0 and recs[0]["hi"] == recs[1]["lo"] and recs[1]["hi"] == len(scan)
    for r in recs:
        assert np.allclose(r["rec"], full, rtol=1e-6, atol=1e-6 * np.max(np.abs(full)))
        assert r["rec"][28] == full[28]                       # inlier counts add up exactly
        assert bytes(r["uid"]) == b"x" * 128                  # the id created on rank 0 reached every rank
    assert np.array_equal(recs[0]["rec"], recs[1]["rec"])     # all ranks solve the same system


def test_round_robin_tiles_match_full(tmp_path):
    """The c5 partition (distributed.interleaved_tiles): every rank linearises its tiles from all over the scan;
    the reduced record is the record of the full scan."""
    world, method = 2, orc.PLANE
    mp.spawn(worker, args=(world, free_port(), method, str(tmp_path), 5), nprocs=world, join=True)
    target = ds.make_urban_slab(30000, seed=3)
    scan = ds.perturb_scan(target, seed=4, num_points=20001)
    tg = orc.build_target(method, target, max_dist=2.0, k=10, voxel_size=1.0)
    T = np.eye(4)
    T[:3, 3] = [0.02, -0.01, 0.03]
    full = record_of(*orc.linearize(tg, T, scan))
    recs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for r in recs:
        assert np.allclose(r["rec"], full, rtol=1e-6, atol=1e-6 * np.max(np.abs(full)))
        assert r["rec"][28] == full[28]
    assert np.array_equal(recs[0]["rec"], recs[1]["rec"])
