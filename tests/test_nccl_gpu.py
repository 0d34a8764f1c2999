"""The path's collective on hardware: two NCCL ranks, one GPU each.

Each rank samples ITS shard of the chains (global chain ids key the Philox streams), counts on its
own GPU with ``out_is_device`` straight into slices of one packed int64 CUDA tensor, and ONE
all-reduce per stage merges the counts over NVLink.  The candidate groups every rank ends up with
must equal, bit for bit, what a single rank computes from all chains.  Needs 2 GPUs
(``gpurun --gpus 2``); skipped cleanly on a one-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CHAINS, N, BURN, SEED = 6, 80, 20, 11


def _data():
    rng = np.random.default_rng(77)
    n, p = 90, 150
    X = rng.standard_normal((n, p))
    for j in range(1, p):
        X[:, j] = 0.9 * X[:, j - 1] + np.sqrt(1 - 0.81) * X[:, j]
    beta = np.zeros(p)
    beta[[10, 11, 60, 100, 140]] = [1.5, -1.0, 0.8, -2.0, 0.6]
    y = X @ beta + rng.standard_normal(n)
    T = 200
    jumps = np.zeros(T)
    jumps[[50, 120]] = [2.5, -2.0]
    Y = np.cumsum(jumps) + 0.7 * rng.standard_normal(T)
    locs = rng.random((240, 5, 2))
    locs[rng.random((240, 5)) < 0.3] = np.nan
    return X, y, Y, locs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_path(rank, world):
    """What every rank (or the single process) computes."""
    import blip_testutil as util
    from pyblip_b200 import LinearSpikeSlab, changepoint, create_groups as cg, create_groups_cts as cts, dist
    X, y, Y, locs = _data()
    lm = LinearSpikeSlab(X=X, y=y)
    lm.sample(N=N, burn=BURN, chains=CHAINS, seed=SEED)
    s0 = dict(dist.STATS)
    seq = util.cg_list(cg.sequential_groups(lm, max_pep=0.5, prenarrow=True, q=0.1))
    s1 = dict(dist.STATS)
    allg = util.cg_list(cg.all_cand_groups(lm, q=0.1, max_pep=0.5, max_size=10))
    s2 = dict(dist.STATS)
    cp = changepoint.ChangepointSpikeSlab(Y)
    cp.sample(N=N, burn=BURN, chains=CHAINS, seed=SEED + 1)
    cps = util.cg_list(changepoint.changepoint_cand_groups(cp, max_pep=0.5))
    fwer = cg.selection_fwer(lm, [[10, 11], [100]])
    # continuous regions: every rank passes its slice of the samples
    lo = locs[rank::world] if world > 1 else np.concatenate([locs[r::2] for r in range(2)])
    peps = cts.grid_peps(lo, grid_sizes=[4, 8, 16], max_pep=0.9)
    diff = lambda a, b, k: a[k] - b[k]   # noqa: E731
    return dict(seq=seq, all=sorted(allg), cps=cps, fwer=fwer, peps=sorted(peps.items()),
                rows=lm.p0s.shape[0], chain_range=lm.chain_range,
                seq_coll=(diff(s1, s0, "all_reduce"), diff(s1, s0, "all_gather"), diff(s1, s0, "device_merges")),
                all_coll=(diff(s2, s1, "all_reduce"), diff(s2, s1, "all_gather"), diff(s2, s1, "device_merges")))


def _worker(rank, world, port, q):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as td
    # deliberately NOT calling torch.cuda.set_device: the library places its buffers itself
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        q.put((rank, _run_path(rank, world)))
    finally:
        td.destroy_process_group()


def test_two_nccl_ranks_equal_one_rank(gpu):
    if gpu < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    want = _run_path(0, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert results[0]["rows"] + results[1]["rows"] == want["rows"] == CHAINS * N
    assert results[0]["chain_range"] == (0, 3) and results[1]["chain_range"] == (3, 6)
    for r in range(2):
        got = results[r]
        # every count of a stage went through ONE NCCL all-reduce of a packed device buffer
        assert got["seq_coll"] == (1, 0, 1), got["seq_coll"]
        assert got["all_coll"] == (3, 1, 3), got["all_coll"]
        assert got["seq"] == want["seq"]
        assert got["all"] == want["all"]
        assert got["cps"] == want["cps"]
        assert got["fwer"] == want["fwer"]
        assert got["peps"] == want["peps"]
