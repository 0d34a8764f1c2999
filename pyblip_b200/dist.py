"""Chain sharding and count merging across the GPUs of one box.

One process per GPU (``torchrun``); chains are independent units (the reference
maps them over a process pool, /root/reference/pyblip/utilities.py:131-137), so
rank r simply runs its slice of the chains with global chain ids as Philox keys.
The only collective on the path is a SUM all-reduce of integer inclusion counts
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # torch missing: single process
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def world():
    """(rank, world_size) of the current process group, (0, 1) if none."""
    d = _dist()
    if d is None:
        return 0, 1
    return d.get_rank(), d.get_world_size()


def shard_chains(chains, rank=None, world_size=None):
    """Contiguous slice [c0, c1) of the global chain ids owned by `rank`.
    The first (chains % world_size) ranks get one extra chain."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(int(chains), int(world_size))
    c0 = rank * base + min(rank, extra)
    c1 = c0 + base + (1 if rank < extra else 0)
    return c0, c1


def allreduce_counts(counts):
    """Sum an int64 numpy array (or a torch tensor, in place) over all ranks."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return counts
    import torch
    if isinstance(counts, torch.Tensor):
        d.all_reduce(counts, op=d.ReduceOp.SUM)
        return counts
    arr = np.ascontiguousarray(counts, dtype=np.int64)
    backend = d.get_backend()
    if backend == "nccl":
        t = torch.from_numpy(arr).cuda()
        d.all_reduce(t, op=d.ReduceOp.SUM)
        return t.cpu().numpy()
    t = torch.from_numpy(arr.copy())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.numpy()


def allreduce_scalar(value):
    return int(allreduce_counts(np.array([int(value)], dtype=np.int64))[0])
