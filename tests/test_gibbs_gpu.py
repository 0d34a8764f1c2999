"""CUDA sampler vs the reference (golden fixtures) and vs the CPU oracle.

Contract (BASELINE.json north_star): driven by an identical injected stream,
indicators are bit-exact and betas / sigma2 / tau2 / p0 agree within 1e-9 relative
for the first 50 sweeps.  All calls go through the C ABI (ctypes).
"""
import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # the tolerance north_star states for floating point results


def _run_device(case, mode, chunk_sweeps=0, gram_refresh=0, inject=True):
    from pyblip_b200 import _lib
    hp = case["hp"]
    cfg = _lib.SpikeSlabConfig(hp["tau2"], int(hp["update_tau2"]), hp["tau2_a0"], hp["tau2_b0"],
                               hp["sigma2"], int(hp["update_sigma2"]), hp["sigma2_a0"],
                               hp["sigma2_b0"], hp["p0"], int(hp["update_p0"]), hp["min_p0"],
                               hp["p0_a0"], hp["p0_b0"])
    design = _lib.Design.upload(case["X"])
    streams = util.stacked_streams(case) if inject else None
    kw = dict(chains=case["chains"], chain_offset=0, n_keep=case["N"], burn=0,
              seed=case["tn_seed"], streams=streams, mode=mode, chunk_sweeps=chunk_sweeps,
              gram_refresh=gram_refresh)
    if case["probit"]:
        smp = _lib.sample_spikeslab(design, cfg, z=case["y"].astype(np.int64), probit=True, **kw)
    else:
        smp = _lib.sample_spikeslab(design, cfg, y=case["y"], probit=False, **kw)
    N, p, C = case["N"], case["X"].shape[1], case["chains"]
    out = dict(betas=smp.dense().reshape(C, N, p))
    p0s, tau2s, sigma2s = smp.hyper()
    out.update(p0s=p0s.reshape(C, N), tau2s=tau2s.reshape(C, N), sigma2s=sigma2s.reshape(C, N))
    if case["probit"]:
        out["y_latent"] = smp.latents().reshape(C, N, case["X"].shape[0])
        out["n_events"] = smp.info.n_events
    out["info"] = smp.info
    # CSR export must agree with the dense materialisation
    indptr, indices, values = smp.csr()
    dense2 = np.zeros((C * N, p))
    for r in range(C * N):
        dense2[r, indices[indptr[r]:indptr[r + 1]]] = values[indptr[r]:indptr[r + 1]]
    assert np.array_equal(dense2, out["betas"].reshape(C * N, p))
    smp.close()
    design.close()
    return out


def _check_against_reference(case, out, tol=RTOL):
    for c, cd in enumerate(case["chains_data"]):
        ref = cd["ref"]
        assert np.array_equal(out["betas"][c] != 0, ref["betas"] != 0), \
            f"{case['name']} chain {c}: indicators differ from the reference"
        for key in ("betas", "p0s", "tau2s", "sigma2s"):
            err = util.rel_err(out[key][c], ref[key])
            assert err <= tol, f"{case['name']} chain {c}: {key} rel err {err:.3e} > {tol}"
        if case["probit"]:
            err = util.rel_err(out["y_latent"][c], ref["y_latent"])
            assert err <= tol, f"{case['name']} chain {c}: latents rel err {err:.3e}"


@pytest.mark.parametrize("name", util.SAMPLER_CASES)
def test_residual_mode_matches_reference(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_RESIDUAL)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", util.LINEAR_CASES)
def test_gram_mode_matches_reference(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_GRAM)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", ["linear_c1", "linear_dense"])
def test_gram_mode_with_refresh_and_small_chunks(gpu, name):
    """Rebuilding q from X'y every 7 sweeps and relaunching every 3 sweeps must
    stay inside the same tolerance (state round-trips through HBM)."""
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_GRAM, chunk_sweeps=3, gram_refresh=7)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", ["linear_c1", "probit_small"])
def test_residual_mode_small_chunks(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_RESIDUAL, chunk_sweeps=7)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", ["linear_c1", "linear_minp0", "linear_odd", "probit_small"])
def test_keyed_philox_stream_matches_oracle(gpu, name):
    """No injection: device Philox / Fisher-Yates / Marsaglia-Tsang / truncated
    normal against the oracle's independent implementation of the stream contract."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    modes = [_lib.MODE_RESIDUAL] if case["probit"] else [_lib.MODE_RESIDUAL, _lib.MODE_GRAM]
    hp = go.hparams(**case["hp"])
    N = case["N"]
    for mode in modes:
        out = _run_device(case, mode, inject=False)
        for c in range(case["chains"]):
            if case["probit"]:
                orc = go.sample_spikeslab(N, case["X"], z=case["y"].astype(np.int64), probit=True,
                                          hp=hp, seed=case["tn_seed"], chain=c)
            else:
                orc = go.sample_spikeslab(N, case["X"], y=case["y"], hp=hp,
                                          seed=case["tn_seed"], chain=c)
            assert np.array_equal(out["betas"][c] != 0, orc["betas"] != 0), (name, mode, c)
            for key in ("betas", "p0s", "tau2s", "sigma2s"):
                err = util.rel_err(out[key][c], orc[key])
                assert err <= RTOL, f"{name} mode {mode} chain {c}: {key} rel err {err:.3e}"
            if case["probit"]:
                assert util.rel_err(out["y_latent"][c], orc["y_latent"]) <= RTOL


def test_chain_offset_gives_same_chains_as_one_launch(gpu):
    """Sharding chains over ranks must not change any chain (Philox key = global id)."""
    from pyblip_b200 import _lib
    case = util.load_sampler_case("linear_c1")
    hp = case["hp"]
    cfg = _lib.SpikeSlabConfig(hp["tau2"], 1, hp["tau2_a0"], hp["tau2_b0"], hp["sigma2"], 1,
                               hp["sigma2_a0"], hp["sigma2_b0"], hp["p0"], 1, hp["min_p0"],
                               hp["p0_a0"], hp["p0_b0"])
    design = _lib.Design.upload(case["X"])
    full = _lib.sample_spikeslab(design, cfg, y=case["y"], chains=5, n_keep=20, burn=5, seed=11,
                                 mode=_lib.MODE_GRAM)
    a = _lib.sample_spikeslab(design, cfg, y=case["y"], chains=2, chain_offset=0, n_keep=20,
                              burn=5, seed=11, mode=_lib.MODE_GRAM)
    b = _lib.sample_spikeslab(design, cfg, y=case["y"], chains=3, chain_offset=2, n_keep=20,
                              burn=5, seed=11, mode=_lib.MODE_GRAM)
    assert np.array_equal(full.dense(), np.concatenate([a.dense(), b.dense()]))
    assert np.array_equal(np.stack(full.hyper()),
                          np.concatenate([np.stack(a.hyper()), np.stack(b.hyper())], axis=1))


def test_gram_build_matches_numpy(gpu):
    from pyblip_b200 import _lib
    rng = np.random.default_rng(5)
    for n, p in [(100, 500), (37, 45), (300, 129), (1000, 257)]:
        X = rng.standard_normal((n, p))
        d = _lib.Design.upload(X)
        d.build_gram()
        G = d.gram()
        ref = X.T @ X
        assert np.max(np.abs(G - ref)) <= 1e-10 * np.max(np.abs(ref)), (n, p)
        assert np.array_equal(G, G.T)
        assert util.rel_err(d.xl2(), (X ** 2).sum(0)) <= 1e-13
        d.close()


def _mc_gate(dev, ref, floor):
    """|mean difference| <= 4 * Monte-Carlo standard error, SE from between-chain variance."""
    se = np.sqrt(dev.var(0, ddof=1) / dev.shape[0] + ref.var(0, ddof=1) / ref.shape[0])
    return np.abs(dev.mean(0) - ref.mean(0)) <= 4 * np.maximum(se, floor)


def test_longrun_pips_within_4_mc_standard_errors_linear(gpu):
    """Long-run posterior summaries of the device sampler (keyed Philox stream) against the
    reference run with its native RNG (tests/golden/longrun_reference.npz): PIPs, posterior
    means and the mean of p0 agree within 4 Monte-Carlo standard errors."""
    import os
    from pyblip_b200 import LinearSpikeSlab
    d = np.load(os.path.join(util.GOLDEN, "longrun_reference.npz"))
    burn, N = int(d["lin_burn"]), int(d["lin_N"])
    chains = d["lin_pips"].shape[0]
    for mode in ("gram", "residual"):
        lm = LinearSpikeSlab(d["lin_X"], d["lin_y"])
        lm.sample(N=N, burn=burn, chains=chains, seed=77, mode=mode)
        b = lm.betas.reshape(chains, N, -1)
        pips = (b != 0).mean(1)
        means = b.mean(1)
        assert _mc_gate(pips, d["lin_pips"], 2e-3).all(), mode
        assert _mc_gate(means, d["lin_means"], 2e-3).all(), mode
        p0 = lm.p0s.reshape(chains, N).mean(1, keepdims=True)
        assert _mc_gate(p0, d["lin_p0"][:, None], 1e-3).all(), mode
        # the true signals are found (reference test_mcmc.py:87-134 style check)
        # the signals the reference finds (its own mean PIP above 0.9) are found here too
        # (reference test_mcmc.py:87-134 style check)
        strong = (d["lin_beta"] != 0) & (d["lin_pips"].mean(0) > 0.9)
        assert strong.any() and pips.mean(0)[strong].min() > 0.9, mode


def test_longrun_pips_within_4_mc_standard_errors_probit(gpu):
    import os
    from pyblip_b200 import ProbitSpikeSlab
    d = np.load(os.path.join(util.GOLDEN, "longrun_reference.npz"))
    burn, N = int(d["pro_burn"]), int(d["pro_N"])
    chains = d["pro_pips"].shape[0]
    pm = ProbitSpikeSlab(d["pro_X"], d["pro_y"])
    pm.sample(N=N, burn=burn, chains=chains, seed=78)
    b = pm.betas.reshape(chains, N, -1)
    assert _mc_gate((b != 0).mean(1), d["pro_pips"], 3e-3).all()
    assert _mc_gate(b.mean(1), d["pro_means"], 3e-3).all()
    Z = pm.Z
    assert Z.shape == (chains * N, d["pro_X"].shape[0])
    # latent signs follow the labels (1 => latent < 0) wherever a row was recorded
    rec = np.any(Z != 0, axis=1)
    assert rec.mean() > 0.5
    lab = d["pro_y"].astype(bool)
    assert np.all(Z[rec][:, lab] < 0) and np.all(Z[rec][:, ~lab] > 0)


def test_conjugate_posterior_on_device(gpu):
    """Reference test_mcmc.py:12-57: hyper-updates off, p0 ~ 0 => conjugate Gaussian posterior."""
    from pyblip_b200 import LinearSpikeSlab
    rng = np.random.default_rng(123)
    n, p, M = 20, 3, 10000
    X = rng.standard_normal((n, p))
    X = X - X.mean(axis=0)
    y = X @ rng.standard_normal(p) + rng.standard_normal(n)
    S11I = np.linalg.inv(np.eye(n) + X @ X.T)
    cond_mean = X.T @ S11I @ y
    cond_var = np.eye(p) - X.T @ S11I @ X
    for mode in ("gram", "residual"):
        lm = LinearSpikeSlab(X=X, y=y, tau2=1, update_tau2=False, sigma2=1, update_sigma2=False,
                             p0=0.000000001, update_p0=False)
        lm.sample(N=M, burn=M, chains=1, seed=5, mode=mode)
        np.testing.assert_array_almost_equal(lm.betas.mean(axis=0), cond_mean, decimal=2)
        np.testing.assert_array_almost_equal(np.cov(lm.betas.T), cond_var, decimal=2)


def test_changepoint_structured_design_equals_dense_design(gpu):
    """The implicit cumulative-sum design (closed-form Gram T - max(a,b)) must reproduce the
    dense design X[i,j] = 1{i >= j} of the reference (changepoint.py:33-35): same keyed
    stream => same indicators, coefficients within 1e-9; and both match the CPU oracle."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import LinearSpikeSlab, changepoint
    rng = np.random.default_rng(9)
    T = 150
    beta = np.zeros(T)
    beta[[30, 80, 120]] = [3.0, -4.0, 2.5]
    Y = np.cumsum(beta) + rng.standard_normal(T)
    X = np.tril(np.ones((T, T)))
    cp = changepoint.ChangepointSpikeSlab(Y)
    cp.sample(N=40, burn=10, chains=3, seed=21)
    lm = LinearSpikeSlab(X, Y)
    lm.sample(N=40, burn=10, chains=3, seed=21, mode="gram")
    assert np.array_equal(cp.betas != 0, lm.betas != 0)
    assert util.rel_err(cp.betas, lm.betas) <= RTOL
    assert util.rel_err(cp.sigma2s, lm.sigma2s) <= RTOL
    for c in range(3):
        orc = go.sample_spikeslab(50, X, y=Y, hp=go.hparams(), seed=21, chain=c)
        assert np.array_equal(cp.betas[c * 40:(c + 1) * 40] != 0, orc["betas"][10:] != 0)
        assert util.rel_err(cp.betas[c * 40:(c + 1) * 40], orc["betas"][10:]) <= RTOL
    from pyblip_b200 import create_groups as cg
    groups = util.cg_list(changepoint.changepoint_cand_groups(cp, max_pep=0.5))
    dense = cp.betas != 0
    dense[:, 0] = 0                                 # changepoint.py:7
    assert groups == util.cg_list(cg.sequential_groups(dense, max_pep=0.5))
    assert len(groups) > 0 and (0,) not in [g for g, _ in groups]


def test_changepoint_block_sampler_on_the_implicit_design(gpu):
    """`detect_changepoints(..., sample_kwargs=dict(bsize=3))` (the reference forwards any sample_kwargs,
    changepoint.py:37-40): the block sampler on the implicit design equals the block sampler on the dense
    design -- same indicators, coefficients within 1e-9 -- and the block oracle."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import LinearSpikeSlab, changepoint
    rng = np.random.default_rng(19)
    T = 140
    beta = np.zeros(T)
    beta[[25, 70, 71, 110]] = [3.0, -2.0, -2.0, 2.5]
    Y = np.cumsum(beta) + rng.standard_normal(T)
    X = np.tril(np.ones((T, T)))
    for bsize, cap in [(2, None), (3, 2), (5, None)]:
        cp = changepoint.ChangepointSpikeSlab(Y)
        cp.sample(N=30, burn=5, chains=2, seed=33, bsize=bsize, max_signals_per_block=cap)
        lm = LinearSpikeSlab(X, Y)
        lm.sample(N=30, burn=5, chains=2, seed=33, mode="gram", bsize=bsize, max_signals_per_block=cap)
        assert np.array_equal(cp.betas != 0, lm.betas != 0), bsize
        # Neighbouring columns of this design differ in ONE row: a block's Gram matrix has condition number
        # of the order of 4 T (asserted below for the first block: > T), and the members of a block are one joint draw --
        # so a coefficient is compared against the largest coefficient of its row (all blocks share the
        # residual), not against itself: a pair (+b, -b') of neighbouring jumps cancels to a small member.
        G = X[:, :bsize].T @ X[:, :bsize]
        assert np.linalg.cond(G) > T
        def scaled(a, b):
            return float(np.max(np.abs(a - b) / np.maximum(np.abs(b).max(1, keepdims=True), 1e-300)))
        e1, e1own = scaled(cp.betas, lm.betas), util.rel_err(cp.betas, lm.betas)
        assert util.rel_err(cp.sigma2s, lm.sigma2s) <= RTOL
        orc = go.sample_spikeslab_multi(35, bsize, X, y=Y, hp=go.hparams(), seed=33, chain=1,
                                        max_signals_per_block=cap or 0)
        assert np.array_equal(cp.betas[30:] != 0, orc["betas"][5:] != 0)
        e2, e2own = scaled(cp.betas[30:], orc["betas"][5:]), util.rel_err(cp.betas[30:], orc["betas"][5:])
        print(f"[changepoint block] bsize {bsize}: vs dense design {e1:.2e} (own magnitude {e1own:.2e}), "
              f"vs oracle {e2:.2e} (own magnitude {e2own:.2e})")
        assert e1 <= RTOL and e2 <= RTOL


def test_api_surface_matches_reference(gpu):
    """Constructor defaults and attributes of the drop-in classes (linear.py:63-80, 165-168)."""
    import inspect
    from pyblip_b200 import LinearSpikeSlab, ProbitSpikeSlab
    sig = inspect.signature(LinearSpikeSlab.__init__)
    want = dict(p0=0.9, p0_a0=1, p0_b0=1, update_p0=True, min_p0=0, sigma2=1, update_sigma2=True,
                sigma2_a0=2, sigma2_b0=1, tau2=1, tau2_a0=2, tau2_b0=1, update_tau2=True)
    for k, v in want.items():
        assert sig.parameters[k].default == v
    ssig = inspect.signature(LinearSpikeSlab.sample)
    assert [ssig.parameters[k].default for k in ("burn", "chains", "num_processes", "bsize",
                                                  "max_signals_per_block")] == [100, 1, 1, 1, None]
    rng = np.random.default_rng(0)
    X = np.asfortranarray(rng.standard_normal((30, 12)))       # non-contiguous input accepted
    y = X[:, 0] * 2 + rng.standard_normal(30)
    lm = LinearSpikeSlab(X, y)
    lm.sample(N=20, burn=5, chains=3, num_processes=4)
    assert lm.betas.shape == (60, 12) and lm.p0s.shape == lm.tau2s.shape == lm.sigma2s.shape == (60,)
    pm = ProbitSpikeSlab(X, (y < 0).astype(float))
    pm.sample(N=10, burn=2, chains=2)
    assert pm.betas.shape == (20, 12) and pm.Z.shape == (20, 30)


def test_truncnorm_values_equal_the_references_own_draws(gpu):
    """a4, value level: the device arithmetic of sample_truncnorm (_truncnorm.pyx:42-107) fed with the
    rand()/RAND_MAX and np.random.randn() sequences the REFERENCE consumed reproduces the reference's
    20 000 recorded values (tests/golden/truncnorm_values.npz, oracle/pin_truncnorm.py): the same
    accept/reject decisions (identical variate consumption per call) and values within 4 ulp
    (CUDA's log/exp/sqrt are not glibc's; the C oracle is bit-identical, tests/test_oracle.py)."""
    import os
    from pyblip_b200 import _lib
    d = np.load(os.path.join(util.GOLDEN, "truncnorm_values.npz"))
    out, uu, zu = _lib.debug_truncnorm(d["mean"], d["var"], d["b"], d["lower"], d["u"], d["u_at"], d["z"], d["z_at"])
    nu = np.diff(np.append(d["u_at"], d["u"].shape[0]))
    nz = np.diff(np.append(d["z_at"], d["z"].shape[0]))
    assert np.array_equal(uu, nu) and np.array_equal(zu, nz), "accept/reject decisions differ"
    ref = d["ref"]
    # error scale: a value mean + scale*z is a difference of terms of that size
    scale = np.maximum(np.abs(ref), np.abs(d["mean"]))
    assert np.max(np.abs(out - ref) / scale) <= 4 * np.finfo(float).eps
    assert np.mean(out == ref) > 0.5
