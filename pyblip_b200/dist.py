"""Chain sharding and count merging across the GPUs of one box.

One process per GPU (``torchrun``); chains are independent units (the reference
maps them over a process pool, /root/reference/pyblip/utilities.py:131-137), so
rank r simply runs its slice of the chains with global chain ids as Philox keys.

The only exchange on the path is the merge of integer inclusion counts.  Every stage of a
counting entry point puts ALL its requests (row count, column counts, window zero-counts,
group any-counts, pair counts ...) into one packed int64 buffer -- the counting kernels write
straight into slices of a CUDA tensor (``out_is_device``) -- and the stage is merged by ONE
SUM all-reduce (NCCL over NVLink on GPUs; gloo on a host buffer in the CPU tests).
``sequential_groups`` / ``changepoint_cand_groups`` are one stage; ``all_cand_groups`` is three
(marginal counts -> per-threshold window and pair counts -> tree-group counts), because each
stage's requests depend on the merged result of the one before.
"""
import os
import sys
import time

import numpy as np

_DEBUG = bool(os.environ.get("BLIPGPU_DIST_DEBUG"))

# number of collectives issued by this module since import (tests assert on it)
STATS = dict(all_reduce=0, all_gather=0, device_merges=0)


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


def _torch_device(device_index):
    import torch
    return torch.device("cuda", int(device_index))


def allreduce_counts(counts, device_index=None):
    """Sum an int64 numpy array (or a torch tensor, in place) over all ranks.  Host arrays under
    NCCL are staged on the GPU ``device_index`` (the library context's device, NOT torch's
    current device: nothing here relies on the caller having called ``torch.cuda.set_device``)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return counts
    import torch
    STATS["all_reduce"] += 1
    if isinstance(counts, torch.Tensor):
        d.all_reduce(counts, op=d.ReduceOp.SUM)
        return counts
    arr = np.ascontiguousarray(counts, dtype=np.int64)
    if d.get_backend() == "nccl":
        if device_index is None:
            from . import _lib
            device_index = _lib.Context.get().device
        t = torch.from_numpy(arr).to(_torch_device(device_index))
        d.all_reduce(t, op=d.ReduceOp.SUM)
        return t.cpu().numpy()
    t = torch.from_numpy(arr.copy())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.numpy()


def allreduce_scalar(value, device_index=None):
    return int(allreduce_counts(np.array([int(value)], dtype=np.int64), device_index)[0])


class CountBatch:
    """All count requests of one stage against one or more sample sets -> one packed int64
    buffer -> ONE all-reduce.

    ``add(dev, kind, *args)`` registers a request and returns its index; ``run()`` returns the
    merged results in request order (int64 arrays shaped like the single-request calls).
    ``dev`` is a ``create_groups._DeviceSamples`` (or the NumPy stand-in of the CPU tests): it
    exposes ``N_local``, ``size_of(kind, *args)``, ``local(kind, *args)`` (host result) and, for
    device-resident sets, ``local_into(kind, device_ptr, *args)`` plus ``device_index``.
    The head of the buffer carries each distinct root set's per-rank row counts, so the global N
    rides along with the first merge instead of costing its own collective."""

    def __init__(self):
        self.reqs = []          # (dev, kind, args, offset, shape)
        self.roots = []         # sample sets whose global N is still unknown
        self.total = 0

    def want_rows(self, dev):
        if dev.N is None and all(r is not dev for r in self.roots):
            self.roots.append(dev)

    def add(self, dev, kind, *args):
        shape = dev.size_of(kind, *args)
        size = int(np.prod(shape)) if len(shape) else 1
        self.reqs.append((dev, kind, args, self.total, shape, size))
        self.total += size
        self.want_rows(dev)
        return len(self.reqs) - 1

    def run(self):
        d = _dist()
        multi = d is not None and d.get_world_size() > 1
        if not multi:
            for r in self.roots:
                r.N = r.N_local
                r.N_per_rank = np.array([r.N_local], dtype=np.int64)
            return [np.asarray(dev.local(kind, *args), dtype=np.int64).reshape(shape)
                    for dev, kind, args, _, shape, _ in self.reqs]
        rank, ws = d.get_rank(), d.get_world_size()
        # every root set gets `world_size` slots and fills its own: the merge yields the row
        # count of every rank (the sum is N; the vector sizes a later all_gather of rows)
        head = len(self.roots) * ws
        total = head + self.total
        on_device = d.get_backend() == "nccl" and all(hasattr(dev, "local_into") for dev, *_ in self.reqs) \
            and all(hasattr(r, "device_index") for r in self.roots)
        if on_device:
            import torch
            index = (self.reqs[0][0] if self.reqs else self.roots[0]).device_index
            # torch.empty launches nothing; every counting call zero-fills its own slice on the
            # library's stream and synchronises that stream before it returns
            t_0 = time.perf_counter()
            buf = torch.empty(max(total, 1), dtype=torch.int64, device=_torch_device(index))
            base = buf.data_ptr()
            for dev, kind, args, off, _, size in self.reqs:
                if size:
                    dev.local_into(kind, base + 8 * (head + off), *args)
            t_1 = time.perf_counter()
            if head:
                hv = np.zeros(head, dtype=np.int64)
                for i, r in enumerate(self.roots):
                    hv[i * ws + rank] = r.N_local
                buf[:head] = torch.from_numpy(hv).to(buf.device)
            STATS["all_reduce"] += 1
            STATS["device_merges"] += 1
            d.all_reduce(buf, op=d.ReduceOp.SUM)
            merged = buf.cpu().numpy()
            if _DEBUG:
                print(f"[dist rank {rank}] stage of {len(self.reqs)} requests, {total} int64: local counts "
                      f"{1e3 * (t_1 - t_0):.1f} ms, all-reduce + read back {1e3 * (time.perf_counter() - t_1):.1f} ms",
                      file=sys.stderr, flush=True)
        else:
            host = np.zeros(max(total, 1), dtype=np.int64)
            for i, r in enumerate(self.roots):
                host[i * ws + rank] = r.N_local
            for dev, kind, args, off, _, size in self.reqs:
                if size:
                    host[head + off:head + off + size] = np.asarray(dev.local(kind, *args), dtype=np.int64).ravel()
            merged = np.asarray(allreduce_counts(host))
        for i, r in enumerate(self.roots):
            r.N_per_rank = merged[i * ws:(i + 1) * ws].copy()
            r.N = int(r.N_per_rank.sum())
        return [merged[head + off:head + off + size].reshape(shape).copy()
                for _, _, _, off, shape, size in self.reqs]


def allgather_rows(words_local, rows_per_rank, device_index=None):
    """Row-wise concatenation over ranks of a C-contiguous uint32 matrix ``[rows_local][cols]``
    (``rows_per_rank`` from an earlier merge): the bit-packed indicator rows that feed the host
    clustering.  ONE padded all_gather."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return words_local
    import torch
    world_size = d.get_world_size()
    cols = words_local.shape[1]
    sizes = np.asarray(rows_per_rank, dtype=np.int64)
    assert sizes.shape[0] == world_size and sizes[d.get_rank()] == words_local.shape[0]
    rmax = int(sizes.max())
    pad = np.zeros((rmax, cols), dtype=np.uint32)
    pad[:words_local.shape[0]] = words_local
    STATS["all_gather"] += 1
    if d.get_backend() == "nccl":
        if device_index is None:
            from . import _lib
            device_index = _lib.Context.get().device
        t = torch.from_numpy(pad.view(np.int32)).to(_torch_device(device_index))
        out = torch.empty((world_size,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        d.all_gather_into_tensor(out, t)
        allw = out.cpu().numpy().view(np.uint32)
    else:
        t = torch.from_numpy(pad.view(np.int32).copy())
        outs = [torch.empty_like(t) for _ in range(world_size)]
        d.all_gather(outs, t)
        allw = np.stack([o.numpy() for o in outs]).view(np.uint32)
    return np.concatenate([allw[r, :int(sizes[r])] for r in range(world_size)], axis=0)


def allgather_ragged_int64(local, device_index=None):
    """``local``: int64 ``[k][len_local]`` (len differs per rank) -> list of per-rank arrays.
    A tiny all-reduce exchanges the lengths, one padded all_gather moves the payload (tensors,
    not pickled objects)."""
    d = _dist()
    local = np.ascontiguousarray(local, dtype=np.int64)
    if d is None or d.get_world_size() == 1:
        return [local]
    import torch
    ws, rank = d.get_world_size(), d.get_rank()
    sizes = np.zeros(ws, dtype=np.int64)
    sizes[rank] = local.shape[1]
    sizes = np.asarray(allreduce_counts(sizes, device_index))
    k, lmax = local.shape[0], max(int(sizes.max()), 1)
    pad = np.zeros((k, lmax), dtype=np.int64)
    pad[:, :local.shape[1]] = local
    STATS["all_gather"] += 1
    if d.get_backend() == "nccl":
        if device_index is None:
            from . import _lib
            device_index = _lib.Context.get().device
        t = torch.from_numpy(pad).to(_torch_device(device_index))
        out = torch.empty((ws, k, lmax), dtype=t.dtype, device=t.device)
        d.all_gather_into_tensor(out, t)
        allv = out.cpu().numpy()
    else:
        t = torch.from_numpy(pad.copy())
        outs = [torch.empty_like(t) for _ in range(ws)]
        d.all_gather(outs, t)
        allv = np.stack([o.numpy() for o in outs])
    return [allv[r, :, :int(sizes[r])] for r in range(ws)]
