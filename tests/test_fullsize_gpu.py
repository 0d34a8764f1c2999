"""Parity at BASELINE.json's full problem sizes, through size-independent properties.

The oracle cannot replay thousands of sweeps at these sizes, so each test pins the CUDA path
with a property that must hold whatever the size:
  * two independent formulations (gram / residual) driven by the same keyed stream produce the
    same indicators and coefficients within 1e-9;
  * a short prefix still matches the CPU oracle;
  * relaunching with another chunk size (state through HBM) changes nothing;
  * counting kernels on the device samples equal NumPy on the dense indicator matrix.
"""
import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _ar1(rng, n, p, rho):
    Z = rng.standard_normal((n, p))
    X = np.empty((n, p))
    X[:, 0] = Z[:, 0]
    s = np.sqrt(1 - rho * rho)
    for j in range(1, p):
        X[:, j] = rho * X[:, j - 1] + s * Z[:, j]
    return X


def _sample(design, cfg, mode, **kw):
    from pyblip_b200 import _lib
    smp = _lib.sample_spikeslab(design, cfg, mode=mode, **kw)
    out = dict(csr=smp.csr(), hyper=np.stack(smp.hyper()), info=smp.info, smp=smp)
    return out


def _csr_equal(a, b, tol=RTOL):
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "indicators differ"
    assert util.rel_err(a[2], b[2]) <= tol


def test_c2_gram_and_residual_agree_and_match_oracle(gpu):
    """configs[1]: n=1000, p=10000, AR(1) rho=0.95 (bench workload), 6 chains x 14 sweeps."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    import bench
    X, y, _ = bench.make_data()
    cfg = LinearSpikeSlab(X, y)._config()
    design = _lib.Design.upload(X)
    kw = dict(y=y, chains=6, chain_offset=100, n_keep=10, burn=4, seed=77)
    g = _sample(design, cfg, _lib.MODE_GRAM, **kw)
    r = _sample(design, cfg, _lib.MODE_RESIDUAL, **kw)
    _csr_equal(g["csr"], r["csr"])
    assert util.rel_err(g["hyper"], r["hyper"]) <= RTOL
    # persistent queue + sub-chunks vs one CTA per chain: 400 chains > resident CTAs
    big = _sample(design, cfg, _lib.MODE_GRAM, y=y, chains=400, chain_offset=0, n_keep=10, burn=4, seed=77)
    dense = big["smp"].dense(row0=100 * 10, row1=106 * 10)
    assert np.array_equal(dense, g["smp"].dense())
    # counting kernels on the device samples vs NumPy on the dense indicators
    bits = g["smp"].bits()
    ind = g["smp"].dense() != 0
    assert np.array_equal(bits.count_columns(), ind.sum(0))
    groups = [[0, 1, 2], [9997, 9998, 9999], list(range(500, 525))]
    assert np.array_equal(bits.count_groups(groups), [ind[:, gr].any(1).sum() for gr in groups])
    bits.close()
    # prefix against the oracle (chain 100, 6 sweeps)
    orc = go.sample_spikeslab(6, X, y=y, hp=go.hparams(), seed=77, chain=100)
    first = _sample(design, cfg, _lib.MODE_GRAM, y=y, chains=1, chain_offset=100, n_keep=6, burn=0, seed=77)
    dev = first["smp"].dense()
    assert np.array_equal(dev != 0, orc["betas"] != 0)
    assert util.rel_err(dev, orc["betas"]) <= RTOL
    for o in (g, r, big, first):
        o["smp"].close()
    design.close()


def test_c3_probit_chunking_invariance_and_oracle_prefix(gpu):
    """configs[2]: probit n=2000, p=5000; 4 chains x 10 sweeps."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    from pyblip_b200.probit import ProbitSpikeSlab
    rng = np.random.default_rng(1003)
    n, p = 2000, 5000
    X = _ar1(rng, n, p, 0.5)
    beta = np.zeros(p)
    beta[rng.choice(p, 25, replace=False)] = rng.standard_normal(25)
    z = (X @ beta + rng.standard_normal(n) < 0).astype(np.int64)
    cfg = ProbitSpikeSlab(X, z.astype(float))._config()
    design = _lib.Design.upload(X)
    kw = dict(z=z, probit=True, chains=4, n_keep=10, burn=0, seed=5)
    a = _sample(design, cfg, _lib.MODE_RESIDUAL, chunk_sweeps=10, **kw)
    b = _sample(design, cfg, _lib.MODE_RESIDUAL, chunk_sweeps=3, **kw)
    _csr_equal(a["csr"], b["csr"], tol=0.0)
    assert np.array_equal(a["hyper"], b["hyper"])
    assert np.array_equal(a["smp"].latents(), b["smp"].latents())
    orc = go.sample_spikeslab(4, X, z=z, probit=True, hp=go.hparams(), seed=5, chain=2)
    dev = a["smp"].dense(row0=20, row1=24)
    assert np.array_equal(dev != 0, orc["betas"] != 0)
    assert util.rel_err(dev, orc["betas"]) <= RTOL
    assert util.latent_err(a["smp"].latents(row0=20, row1=24), orc["y_latent"]) <= RTOL
    a["smp"].close(); b["smp"].close(); design.close()


def test_c4_changepoint_full_length_counts(gpu):
    """configs[3]: T=20000 cumulative-sum design (never materialised), 3 chains x 8 sweeps, then the
    sequential-group counts of the device samples against NumPy prefix sums (create_groups.py:321-332)."""
    from oracle import counting_oracle as co
    from pyblip_b200 import changepoint
    rng = np.random.default_rng(1004)
    T = 20000
    beta = np.zeros(T)
    loc = rng.choice(T, 40, replace=False)
    beta[loc] = np.sign(rng.standard_normal(40)) * np.maximum(np.abs(rng.normal(0, np.sqrt(5), 40)), 0.1)
    Y = np.cumsum(beta) + rng.standard_normal(T)
    cp = changepoint.ChangepointSpikeSlab(Y)
    cp.sample(N=8, burn=2, chains=3, seed=31)
    cp2 = changepoint.ChangepointSpikeSlab(Y)
    cp2.sample(N=8, burn=2, chains=3, seed=31, chunk_sweeps=3)
    assert np.array_equal(cp.betas, cp2.betas)
    assert np.isfinite(cp.sigma2s).all() and (cp.sigma2s > 0).all()
    ind = cp.betas != 0
    bits = cp.samples_handle.bits()
    assert np.array_equal(bits.count_columns(), ind.sum(0))
    assert np.array_equal(bits.count_sequential(25), co.sequential_zero_counts(ind, 25))
    bits.close()
    # the largest jumps are found within a few positions after 10 sweeps
    top = np.argsort(-np.abs(beta))[:5]
    hits = [ind[:, max(0, t - 3):t + 4].any(1).mean() for t in top]
    assert min(hits) > 0.5, hits


def test_c5_global_q_gram_kernel_equals_residual(gpu):
    """configs[4] scale: p=50000, n=5000 -- q (400 KB) does not fit shared memory, so the gram kernel
    keeps it in global memory; 3 chains x 5 sweeps against the residual kernel."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    rng = np.random.default_rng(1005)
    n, p, blk = 5000, 50000, 100
    X = np.empty((n, p))
    for b0 in range(0, p, blk):                      # LD-like blocks: AR(1) inside, independent across
        rho = min(rng.beta(5, 1), 0.99)
        X[:, b0:b0 + blk] = _ar1(rng, n, blk, rho)
    beta = np.zeros(p)
    beta[rng.choice(p, 50, replace=False)] = rng.standard_normal(50)
    y = X @ beta + rng.standard_normal(n)
    cfg = LinearSpikeSlab(X, y)._config()
    design = _lib.Design.upload(X)
    kw = dict(y=y, chains=3, n_keep=5, burn=0, seed=9)
    g = _sample(design, cfg, _lib.MODE_GRAM, **kw)
    r = _sample(design, cfg, _lib.MODE_RESIDUAL, **kw)
    # The first sweeps from beta = 0 activate ~5600 coordinates, more than there are observations,
    # so the two accumulation orders (q updated through 5600 Gram columns vs residual dots) differ by
    # a little more than in the stationary regime: element-wise 5e-9 here, indicators still identical.
    _csr_equal(g["csr"], r["csr"], tol=5e-9)
    assert util.rel_err(g["hyper"], r["hyper"]) <= RTOL
    g["smp"].close(); r["smp"].close(); design.close()


def test_nprior_p2000_matches_oracle_prefix(gpu):
    """Neuronized prior at p=2000 (blocked Cholesky path is reached by some chains during burn-in):
    chain 0, first 4 sweeps against the oracle under the keyed stream."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    import bench
    X, _, _ = bench.make_data()
    Xn = np.ascontiguousarray(X[:, :2000])
    rng = np.random.default_rng(7)
    b = np.zeros(2000)
    b[rng.choice(2000, 20, replace=False)] = rng.standard_normal(20)
    yn = Xn @ b + rng.standard_normal(Xn.shape[0])
    kw = dict(tauw2=1.0, p0_init=1 - 1 / 2000, min_p0=1e-10, update_p0=1, sigma_a0=5.0, sigma_b0=1.0,
              sigma_prior_type=0, joint_sample_W=1, group_alpha_update=1)
    cfg = _lib.NPriorConfig(kw["tauw2"], kw["p0_init"], kw["min_p0"], kw["update_p0"], kw["sigma_a0"],
                            kw["sigma_b0"], kw["sigma_prior_type"], kw["joint_sample_W"],
                            kw["group_alpha_update"], 0)
    design = _lib.Design.upload(Xn)
    smp = _lib.sample_nprior(design, cfg, yn, chains=2, n_keep=4, burn=0, seed=1)
    out = {k: smp.get(k) for k in ("alphas", "ws", "betas", "sigma2s", "p0s")}
    smp.close(); design.close()
    orc = go.nprior_sample(4, Xn, yn, seed=1, chain=0, **kw)
    assert np.array_equal(out["betas"][:4] != 0, orc["betas"] != 0)
    for k in ("alphas", "ws", "sigma2s", "p0s"):
        assert util.rel_err(out[k][:4], orc[k]) <= 1e-8, k


def test_c2_queue_scheduling_is_deterministic(gpu):
    """The persistent work queue hands (chain, 8-sweep) items to whichever CTA is free, so two runs
    schedule differently; a race on shared state would show up as run-to-run differences.  400 chains
    x 40 sweeps twice with one seed, and once as two half-size calls with chain offsets: identical
    indicators, coefficients and hyper-parameters, bit for bit."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    import bench
    X, y, _ = bench.make_data()
    cfg = LinearSpikeSlab(X, y)._config()
    design = _lib.Design.upload(X)
    kw = dict(y=y, n_keep=30, burn=10, seed=123)
    a = _sample(design, cfg, _lib.MODE_GRAM, chains=400, chain_offset=0, **kw)
    b = _sample(design, cfg, _lib.MODE_GRAM, chains=400, chain_offset=0, **kw)
    for k in range(3):
        assert np.array_equal(a["csr"][k], b["csr"][k])
    assert np.array_equal(a["hyper"], b["hyper"])
    c1 = _sample(design, cfg, _lib.MODE_GRAM, chains=150, chain_offset=0, **kw)       # one CTA per chain
    c2 = _sample(design, cfg, _lib.MODE_GRAM, chains=250, chain_offset=150, **kw)
    n1 = int(c1["csr"][0][-1])
    assert np.array_equal(a["csr"][0][:150 * 30 + 1], c1["csr"][0])
    assert np.array_equal(a["csr"][1][:n1], c1["csr"][1][:n1]) and np.array_equal(a["csr"][2][:n1], c1["csr"][2][:n1])
    assert np.array_equal(a["csr"][2][n1:int(a["csr"][0][-1])], c2["csr"][2][:int(c2["csr"][0][-1])])
    assert np.array_equal(a["hyper"], np.concatenate([c1["hyper"], c2["hyper"]], axis=1))
    for o in (a, b, c1, c2):
        o["smp"].close()
    design.close()
