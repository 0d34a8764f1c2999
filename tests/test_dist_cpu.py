"""N > 1 path on CPU: two gloo ranks each hold half of the chains' samples; integer counts
are merged with one SUM all-reduce and the resulting candidate groups equal the
single-process result on the concatenated samples."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as td
    td.init_process_group("gloo", rank=rank, world_size=world)
    import blip_testutil as util
    import fake_device
    from pyblip_b200 import create_groups as cg, dist
    fake_device.install(None)
    samples = util.load_counting_case("sparse_a")["samples"]
    chains = 5                                   # rows = chains x sweeps, chain-major
    N = samples.shape[0] // chains
    samples = samples[:chains * N]
    c0, c1 = dist.shard_chains(chains)
    local = samples[c0 * N:c1 * N]
    assert dist.world() == (rank, world)
    n0 = dict(dist.STATS)
    seq = util.cg_list(cg.sequential_groups(local, max_pep=0.5))
    n1 = dict(dist.STATS)
    allg = util.cg_list(cg.all_cand_groups(local, q=0.1, max_pep=0.5, max_size=12))
    n2 = dict(dist.STATS)
    out = {
        "seq": seq, "all": allg, "n_local": local.shape[0],
        # collectives: sequential_groups is ONE all-reduce; all_cand_groups is three stages
        # (marginals -> windows of every threshold -> tree groups of every threshold) plus one
        # all_gather of bit-packed rows for the host clustering of these small sets
        "seq_collectives": (n1["all_reduce"] - n0["all_reduce"], n1["all_gather"] - n0["all_gather"]),
        "all_collectives": (n2["all_reduce"] - n1["all_reduce"], n2["all_gather"] - n1["all_gather"]),
    }
    q.put((rank, out))
    td.destroy_process_group()


def test_two_rank_counting_matches_single_process(monkeypatch):
    import blip_testutil as util
    import fake_device
    from pyblip_b200 import create_groups as cg
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fake_device.install(monkeypatch)
    samples = util.load_counting_case("sparse_a")["samples"]
    N = samples.shape[0] // 5
    samples = samples[:5 * N]
    want_seq = util.cg_list(cg.sequential_groups(samples, max_pep=0.5))
    want_all = util.cg_list(cg.all_cand_groups(samples, q=0.1, max_pep=0.5, max_size=12))
    assert results[0]["n_local"] + results[1]["n_local"] == samples.shape[0]
    assert results[0]["n_local"] != results[1]["n_local"]       # 3 chains vs 2 chains
    for r in range(world):
        assert results[r]["seq_collectives"] == (1, 0), results[r]["seq_collectives"]
        assert results[r]["all_collectives"] == (3, 1), results[r]["all_collectives"]
        assert results[r]["seq"] == want_seq
        assert sorted(results[r]["all"]) == sorted(want_all)
