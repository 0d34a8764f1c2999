"""The CPU oracle against the reference's recorded outputs (tests/golden/*.npz).

The fixtures were produced by oracle/pin_against_reference.py from the reference's own
compiled Cython kernel running with its native random sources; replaying the recorded
stream through the oracle must give identical indicators and <= 1e-11 relative agreement.
"""
import numpy as np
import pytest

import blip_testutil as util
from oracle import gibbs_oracle as go


@pytest.mark.parametrize("name", util.SAMPLER_CASES)
def test_oracle_replays_reference(name):
    case = util.load_sampler_case(name)
    hp = go.hparams(**case["hp"])
    for c, cd in enumerate(case["chains_data"]):
        st = cd["streams"]
        if case["probit"]:
            out = go.sample_spikeslab(case["N"], case["X"], z=case["y"].astype(np.int64),
                                      probit=True, hp=hp, seed=case["tn_seed"], chain=c, **st)
        else:
            out = go.sample_spikeslab(case["N"], case["X"], y=case["y"], hp=hp,
                                      seed=case["tn_seed"], chain=c, **st)
        ref = cd["ref"]
        assert np.array_equal(out["betas"] != 0, ref["betas"] != 0)
        for key in ("betas", "p0s", "tau2s", "sigma2s"):
            assert util.rel_err(out[key], ref[key]) <= 1e-11, (name, c, key)
        if case["probit"]:
            assert util.rel_err(out["y_latent"], ref["y_latent"]) <= 1e-11


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    assert go.philox(0, 0, 0, 0, 0, 0).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = go.philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert out.tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = go.philox(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)
    assert out.tolist() == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_keyed_permutation_is_a_permutation_and_depends_on_key():
    a = go.permutation(5, 0, 0, 1000)
    assert sorted(a.tolist()) == list(range(1000))
    assert not np.array_equal(a, go.permutation(5, 1, 0, 1000))
    assert not np.array_equal(a, go.permutation(5, 0, 1, 1000))
    assert np.array_equal(a, go.permutation(5, 0, 0, 1000))


def test_keyed_gamma_beta_laws():
    import scipy.stats as st
    lib = go.lib()
    for shape in (0.4, 1.0, 2.5, 51.0, 502.0):
        x = np.array([lib.orc_gamma(9, 1, 0, i, shape) for i in range(20000)])
        assert st.kstest(x, st.gamma(shape).cdf).pvalue > 1e-3, shape
    x = np.array([lib.orc_beta(9, 2, 0, i, 451.0, 51.0) for i in range(20000)])
    assert st.kstest(x, st.beta(451.0, 51.0).cdf).pvalue > 1e-3


def test_truncnorm_law_and_support():
    import scipy.stats as st
    lib = go.lib()
    for mean, lower in [(0.0, 0), (1.3, 1), (-2.0, 0), (2.5, 1)]:
        x = np.array([lib.orc_truncnorm(7, 3, e, 0, mean, 1.5, 0.0, lower) for e in range(20000)])
        sd = np.sqrt(1.5)
        if lower:
            assert np.all(x < 0)
            dist = st.truncnorm(-np.inf, (0 - mean) / sd, loc=mean, scale=sd)
        else:
            assert np.all(x > 0)
            dist = st.truncnorm((0 - mean) / sd, np.inf, loc=mean, scale=sd)
        assert st.kstest(x, dist.cdf).pvalue > 1e-3


def test_oracle_conjugate_posterior():
    """Reference test_mcmc.py:12-57 restated for the oracle: with hyper-updates off and
    p0 ~ 0 the chain samples the conjugate Gaussian posterior."""
    rng = np.random.default_rng(123)
    n, p, M = 20, 3, 6000
    X = rng.standard_normal((n, p))
    X = X - X.mean(axis=0)
    beta = rng.standard_normal(p)
    y = X @ beta + rng.standard_normal(n)
    hp = go.hparams(tau2=1, update_tau2=False, sigma2=1, update_sigma2=False, p0=1e-9,
                    update_p0=False)
    out = go.sample_spikeslab(2 * M, X, y=y, hp=hp, seed=3, chain=0)
    S11I = np.linalg.inv(np.eye(n) + X @ X.T)
    cond_mean = X.T @ S11I @ y
    cond_var = np.eye(p) - X.T @ S11I @ X
    b = out["betas"][M:]
    np.testing.assert_allclose(b.mean(0), cond_mean, atol=0.03)
    np.testing.assert_allclose(np.cov(b.T), cond_var, atol=0.03)


@pytest.mark.parametrize("name", ["nprior_joint_group", "nprior_plain", "nprior_sigprior",
                                  "nprior_fixedp0"])
def test_nprior_oracle_replays_reference(name):
    """oracle/nprior_oracle.c against the reference's recorded run (oracle/pin_nprior.py)."""
    import os
    d = np.load(os.path.join(util.GOLDEN, name + ".npz"))
    kw = {k[3:]: float(d[k]) for k in d.files if k.startswith("kw_")}
    for k in ("update_p0", "sigma_prior_type", "joint_sample_W", "group_alpha_update"):
        kw[k] = int(kw[k])
    st = {k[3:]: d[k] for k in d.files if k.startswith("st_")}
    out = go.nprior_sample(int(d["N"]), d["X"], d["y"], seed=int(d["seed"]), chain=0, streams=st, **kw)
    assert np.array_equal(out["betas"] != 0, d["ref_betas"] != 0)
    for k in ("alphas", "ws", "betas", "sigma2s", "alpha0s", "p0s"):
        assert util.rel_err(out[k], d["ref_" + k]) <= 1e-9, (name, k)


@pytest.mark.parametrize("name", util.MULTI_CASES)
def test_block_oracle_replays_reference(name):
    """oracle/multi_oracle.c against the reference's recorded `_sample_spikeslab_multi` runs
    (oracle/pin_multi.py): identical indicators, 1e-9 on coefficients and hyper-parameters."""
    case = util.load_multi_case(name)
    probit = case["probit"]
    for c, cd in enumerate(case["chains_data"]):
        st = dict(cd["streams"])
        out = go.sample_spikeslab_multi(
            case["N"], case["bsize"], case["X"], y=None if probit else case["y"],
            z=case["y"].astype(np.int64) if probit else None, probit=probit,
            hp=go.hparams(**case["hp"]), max_signals_per_block=case["max_signals"],
            seed=case["tn_seed"], chain=c, **st)
        ref = cd["ref"]
        assert np.array_equal(out["betas"] != 0, ref["betas"] != 0), (name, c)
        for k in ("betas", "p0s", "tau2s", "sigma2s"):
            assert util.rel_err(out[k], ref[k]) <= 1e-9, (name, c, k)
        if probit:
            assert util.latent_err(out["y_latent"], ref["y_latent"]) <= 1e-9, (name, c)


def test_block_oracle_partition_and_models():
    """Block table of the keyed stream: ceil(p/bsize) blocks of consecutive wrapping indices
    (_linear_multi.pyx:298,395-402); models in itertools.combinations order (:405-409)."""
    import itertools
    assert [go.multi_nblock(p, b) for p, b in ((10, 3), (10, 4), (8, 4), (9, 3), (500, 5))] == [4, 3, 2, 3, 100]
    for bsize, cap in ((3, 0), (5, 0), (5, 2), (4, 1)):
        want = [sum(1 << j for j in comb) for m in range(1, (cap or bsize) + 1)
                for comb in itertools.combinations(range(bsize), m)]
        assert list(go.multi_models(bsize, cap)) == want


def test_keyed_permutation_is_statistically_uniform():
    """The point-wise Feistel visiting order: every coordinate is equally likely at
    every position, consecutive positions are not correlated, one fixed point per permutation on average."""
    import scipy.stats as st
    p, N = 500, 6000
    P = np.stack([go.permutation(11, 3, s, p) for s in range(N)])
    assert all(sorted(go.permutation(5, 0, 0, q)) == list(range(q)) for q in (1, 2, 3, 17, 1000, 70001))
    pvals = []
    for col in (0, 1, p // 2, p - 1):                       # coordinate at a fixed position
        h = np.bincount(P[:, col], minlength=p)
        pvals.append(st.chi2.sf(((h - N / p) ** 2 / (N / p)).sum(), p - 1))
    inv = np.argsort(P, axis=1)
    for j in (0, 7, p - 1):                                 # position of a fixed coordinate
        h = np.bincount(inv[:, j], minlength=p)
        pvals.append(st.chi2.sf(((h - N / p) ** 2 / (N / p)).sum(), p - 1))
    assert min(pvals) > 1e-4, pvals
    assert abs(np.corrcoef(P[:, 0], P[:, 1])[0, 1]) < 0.06
    assert abs(np.corrcoef(P[:-1, 0], P[1:, 0])[0, 1]) < 0.06
    adj = (np.abs(np.diff(P, axis=1)) == 1).mean()
    assert abs(adj - 2 / p) < 0.15 * 2 / p
    assert abs((P == np.arange(p)).sum(1).mean() - 1.0) < 0.08


def test_truncnorm_arithmetic_reproduces_the_references_values_bit_for_bit():
    """a4, value level: the oracle's truncated-normal arithmetic on the uniform / normal sequences the
    reference's compiled `sample_truncnorm` consumed (libc rand() after srand(s), NumPy legacy randn
    after seed(s)) gives the reference's recorded values exactly (oracle/pin_truncnorm.py wrote the
    fixture from oracle/_ref)."""
    from oracle import pin_truncnorm as pt
    import os
    d = np.load(os.path.join(util.GOLDEN, "truncnorm_values.npz"))
    out, used, u_at, z_at = pt.injected(d["mean"], d["var"], d["b"], d["lower"], d["u"], d["z"])
    assert np.array_equal(out, d["ref"])
    assert used[0] == d["u"].shape[0] and used[1] == d["z"].shape[0]
    assert np.array_equal(u_at, d["u_at"]) and np.array_equal(z_at, d["z_at"])
    a = (d["b"] - d["mean"]) / np.sqrt(d["var"])
    a = np.where(d["lower"] == 1, -a, a)
    # every branch of _truncnorm.pyx:57-86 is in the fixture
    assert (a >= 0.15).sum() > 1000 and ((a >= 0) & (a < 0.15)).sum() > 1000 and (a < 0).sum() > 1000
