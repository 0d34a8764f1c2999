"""Randomised shapes through the counting kernels and the block sampler against the CPU oracle:
ragged N and p, all densities, every window size -- the corners of the sparse-tile, column-major
and window-scoring paths that the fixed fixtures do not pin."""
import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu


def test_counting_kernels_on_random_shapes(gpu):
    from oracle import counting_oracle as co
    from pyblip_b200 import _lib
    rng = np.random.default_rng(2024)
    for case in range(60):
        N = int(rng.integers(1, 3000))
        p = int(rng.integers(1, 300))
        dens = float(rng.choice([0.0, 0.001, 0.02, 0.3, 0.9]))
        S = rng.random((N, p)) < dens
        if p > 3 and case % 3 == 0:
            S[:, rng.integers(0, p, size=3)] = True               # a few dense columns among sparse ones
        max_size = int(rng.integers(1, 41))
        bits = _lib.Bits.from_dense(S)
        assert np.array_equal(bits.count_columns(), co.column_counts(S)), (case, N, p)
        got = bits.count_sequential(max_size)                     # int64[max_size, p]; the oracle clamps to p rows
        want = co.sequential_zero_counts(S, max_size)
        assert np.array_equal(got[:want.shape[0]], want), (case, N, p, dens, max_size)
        assert not got[want.shape[0]:].any(), (case, N, p, max_size)  # windows wider than p hold no group
        ng = int(rng.integers(1, 80))
        groups = [sorted(rng.choice(p, size=int(rng.integers(1, min(p, 12) + 1)), replace=False).tolist())
                  for _ in range(ng)]
        want = np.array([np.any(S[:, g], axis=1).sum() for g in groups], dtype=np.int64)
        assert np.array_equal(bits.count_groups(groups), want), (case, N, p, ng)
        fd = np.zeros(N, dtype=bool)
        for g in groups[:5]:
            fd |= np.all(S[:, g] == 0, axis=1)
        assert bits.count_false_rows(groups[:5]) == int(fd.sum()), (case, N, p)
        bits.close()


def test_block_sampler_on_random_problems(gpu):
    """Keyed stream, both forms, block sizes 2..6 with and without a cap on the signals per block."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    rng = np.random.default_rng(77)
    for case in range(8):
        n = int(rng.integers(20, 90))
        p = int(rng.integers(7, 60))
        bsize = int(rng.integers(2, 7))
        cap = int(rng.integers(0, bsize + 1))
        X = rng.standard_normal((n, p))
        X[:, 1:] += 0.7 * X[:, :-1]                               # correlated neighbours
        beta = np.zeros(p)
        beta[rng.choice(p, size=min(4, p), replace=False)] = 2.0 * rng.standard_normal(min(4, p))
        y = X @ beta + rng.standard_normal(n)
        N, seed = 12, int(rng.integers(1, 1 << 30))
        cfg = LinearSpikeSlab(X, y)._config()
        hp = go.hparams()
        design = _lib.Design.upload(X)
        for mode in (_lib.MODE_RESIDUAL, _lib.MODE_GRAM):
            smp = _lib.sample_spikeslab(design, cfg, y=y, chains=2, n_keep=N, burn=0, seed=seed, mode=mode,
                                        bsize=bsize, max_signals_per_block=cap)
            betas = smp.dense().reshape(2, N, p)
            p0s, tau2s, sigma2s = (a.reshape(2, N) for a in smp.hyper())
            smp.close()
            for c in range(2):
                ref = go.sample_spikeslab_multi(N, bsize, X, y=y, probit=False, hp=hp, max_signals_per_block=cap,
                                                seed=seed, chain=c)
                assert np.array_equal(betas[c] != 0, ref["betas"] != 0), (case, mode, c, n, p, bsize, cap)
                for got, key in ((betas[c], "betas"), (p0s[c], "p0s"), (tau2s[c], "tau2s"), (sigma2s[c], "sigma2s")):
                    assert util.rel_err(got, ref[key]) <= 1e-9, (case, mode, c, key)
        design.close()
