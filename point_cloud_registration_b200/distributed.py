"""Multi-GPU scan sharding: one process per GPU, the target structure
replicated, the scan tile-sharded, one 29-double NCCL all-reduce per iteration inside
libpcr_b200.so.  ``torch.distributed`` is used only as the rendezvous that ships the NCCL
unique id (any backend, e.g. gloo)."""
import numpy as np


def shard_bounds(n, rank, world_size):
    """Contiguous tile [lo, hi) of an n-point scan owned by ``rank``; tile sizes differ by at
    most one point and cover the scan exactly."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def interleaved_tiles(n, rank, world_size, tiles_per_rank=16):
    """Index ranges [(lo, hi), ...] of a spatially ordered n-point scan owned by ``rank``: the scan is cut into
    ``world_size * tiles_per_rank`` contiguous tiles dealt round-robin.  Every tile stays a compact region (its
    queries share list cells), and every rank gets tiles from all over the scene -- the cost of a tile varies with
    its distance from the rotation centre, and one contiguous range per rank left the rim ranks 30 % behind in the
    first iterations (profiles/r2_notes.md section C)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    k_total = int(world_size) * max(1, int(tiles_per_rank))
    tiles = [shard_bounds(n, k, k_total) for k in range(rank, k_total, int(world_size))]
    return [(lo, hi) for lo, hi in tiles if hi > lo]


def exchange_unique_id(make_id, rank, group=None):
    """Rank 0 creates the 128-byte NCCL id with ``make_id()``; every rank returns it."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def attach(registration, group=None):
    """Attach a registration object (target already set) to the default torch.distributed
    process group: NCCL communicator over all ranks, scan sharded by rank."""
    import torch.distributed as dist
    from ._lib import Context
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = exchange_unique_id(Context.comm_unique_id, rank, group)
    registration.attach_communicator(rank, world, uid)
    return rank, world


def allreduce_record_host(rec, group=None):
    """Host-side sum of a 29-double record over the process group (any backend).  Used by the
    gloo tests of the sharding logic and available as a debugging path; the product path
    reduces on the device with NCCL."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(rec, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
