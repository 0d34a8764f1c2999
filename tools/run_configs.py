"""BASELINE.json configs[2..4] at full problem size on one GPU: wall-clock throughput of the
reference-facing API (sampler) and of the candidate-group counting on the device samples.
configs[1] is bench.py; configs[0] (README quick-start) is covered by the tests.

    python tools/run_configs.py [c3] [c4] [c5]  > profiles/r01_configs.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyblip_b200 import ProbitSpikeSlab, LinearSpikeSlab, changepoint, create_groups  # noqa: E402

which = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5"}


def ar1(rng, n, p, rho):
    Z = rng.standard_normal((n, p))
    X = np.empty((n, p))
    X[:, 0] = Z[:, 0]
    s = np.sqrt(1 - rho * rho)
    for j in range(1, p):
        X[:, j] = rho * X[:, j - 1] + s * Z[:, j]
    return X


def report(name, chains, p, sweeps, wall, lm, extra=""):
    i = lm.samples_handle.info
    print(f"{name}: {chains} chains x {sweeps} sweeps x p={p}: wall {wall:.2f} s (device {i.total_ms / 1e3:.2f} s, "
          f"sweep kernels {i.sweep_ms / 1e3:.2f} s) = {chains * p * sweeps / wall:.3e} coordinate updates/s; "
          f"changes/chain-sweep {i.n_changes / i.n_sweeps_total:.1f}; nnz/row {i.nnz / i.rows:.1f} {extra}", flush=True)


if "c2" in which:      # bench workload, then all_cand_groups on the device samples
    import bench
    X, y, beta = bench.make_data()
    lm = LinearSpikeSlab(X, y)
    t0 = time.perf_counter()
    lm.sample(N=5000, burn=100, chains=512, seed=1)
    wall = time.perf_counter() - t0
    report("C2 linear n=1000 p=10000", 512, 10000, 5100, wall, lm)
    t0 = time.perf_counter()
    groups = create_groups.all_cand_groups(lm.samples_handle, X=None, q=0.1,
                                           prefilter_thresholds=[0.01, 0.02, 0.03, 0.05, 0.1, 0.2])
    dt = time.perf_counter() - t0
    true = set(np.flatnonzero(beta).tolist())
    hit = sum(1 for j in true if any(j in g.group for g in groups))
    print(f"   all_cand_groups on 2 560 000 x 10 000 device indicators (thresholds 0.01..0.2; correlation from device "
          f"pair counts): {len(groups)} groups in {dt:.2f} s; {hit}/{len(true)} true signals covered", flush=True)
    lm._release(); lm._design.close()

if "c3" in which:      # probit n=2000 p=5000, N=2000, 256 chains
    rng = np.random.default_rng(1003)
    n, p = 2000, 5000
    X = ar1(rng, n, p, 0.95)        # AR(1) rho = 0.95
    beta = np.zeros(p)
    beta[rng.choice(p, 25, replace=False)] = rng.standard_normal(25)
    z = (X @ beta + rng.standard_normal(n) < 0).astype(float)
    pm = ProbitSpikeSlab(X, z)
    t0 = time.perf_counter()
    pm.sample(N=2000, burn=100, chains=256, seed=1)
    wall = time.perf_counter() - t0
    report("C3 probit n=2000 p=5000", 256, p, 2100, wall, pm)
    t0 = time.perf_counter()
    groups = create_groups.all_cand_groups(pm.samples_handle, X=None, q=0.1, prefilter_thresholds=[0.05, 0.1, 0.2])
    print(f"   all_cand_groups on the device samples (thresholds 0.05, 0.1, 0.2): {len(groups)} groups in "
          f"{time.perf_counter() - t0:.2f} s", flush=True)
    pm._release(); pm._design.close()

if "c4" in which:      # change points T=20000, N=2000, 128 chains + sequential_groups
    rng = np.random.default_rng(1004)
    T = 20000
    beta = np.zeros(T)
    loc = rng.choice(T, 40, replace=False)
    beta[loc] = np.sign(rng.standard_normal(40)) * np.maximum(np.abs(rng.normal(0, np.sqrt(5), 40)), 0.1)
    Y = np.cumsum(beta) + rng.standard_normal(T)
    cp = changepoint.ChangepointSpikeSlab(Y)
    t0 = time.perf_counter()
    cp.sample(N=2000, burn=100, chains=128, seed=2)
    wall = time.perf_counter() - t0
    report("C4 changepoint T=20000", 128, T, 2100, wall, cp)
    t0 = time.perf_counter()
    groups = changepoint.changepoint_cand_groups(cp, max_size=25, max_pep=0.25)
    dt = time.perf_counter() - t0
    found = sum(1 for t in loc if any(t in g.group for g in groups))
    print(f"   changepoint_cand_groups (256000 x 20000 indicators on the device): {len(groups)} groups in {dt:.2f} s; "
          f"{found}/40 true change points covered by a group", flush=True)
    cp._release(); cp._design.close()

if "c5" in which:      # fine-mapping scale p=50000 n=5000, 1024 chains (N=1000 assumed in SURVEY; here N=200)
    rng = np.random.default_rng(1005)
    n, p, blk = 5000, 50000, 100
    X = np.empty((n, p))
    for b0 in range(0, p, blk):
        X[:, b0:b0 + blk] = ar1(rng, n, blk, min(rng.beta(5, 1), 0.99))
    beta = np.zeros(p)
    beta[rng.choice(p, 50, replace=False)] = rng.standard_normal(50)
    y = X @ beta + rng.standard_normal(n)
    lm = LinearSpikeSlab(X, y)
    t0 = time.perf_counter()
    lm.sample(N=200, burn=50, chains=1024, seed=3)
    wall = time.perf_counter() - t0
    report("C5 fine-mapping n=5000 p=50000 (gram form, q in global memory)", 1024, p, 250, wall, lm,
           extra="[incl. H2D of X (2 GB) and the 20 GB Gram build]")
    t0 = time.perf_counter()
    bits = lm.samples_handle.bits()
    cc = bits.count_columns()
    seq = bits.count_sequential(25)
    bits.close()
    print(f"   count_columns + count_sequential(25) on 204800 x 50000 indicators: {time.perf_counter() - t0:.3f} s; "
          f"max marginal PIP {cc.max() / lm.samples_handle.info.rows:.3f}", flush=True)
    lm._release(); lm._design.close()
